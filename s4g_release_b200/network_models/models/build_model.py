"""Model factory — reference network_models/models/build_model.py:13-31: ``cfg.MODEL.TYPE`` -> ``(net, loss, metric)``.

``cfg`` is the reference's yacs-style node (attribute access; yacs itself is not needed — any object tree with the
same attribute names works).  The PointNet++ family of the hot path is served; the two image / PointNet baselines of
the paper's comparison (``GPD``: LeNet on 60x60 projections, ``PointNetGPD``) never touch the pn2_ext operators and are
out of this build's scope: asking for them says so instead of failing with an import error."""
from .EdgePointNet2Down import build_edgepointnet2down
from .EdgePointNet2DownUp import build_edgepointnet2downup
from .PointNet2 import build_pointnet2
from .PointNet2_local import build_pointnet2_local
from .PointNet2_tcls import build_pointnet2_cls

_BUILDERS = {
    "PN2": build_pointnet2,
    "PN2_CLS": build_pointnet2_cls,
    "PN2_LOCAL": build_pointnet2_local,
    "EDGEPN2D": build_edgepointnet2down,
    "EDGEPN2DU": build_edgepointnet2downup,
}
_OUT_OF_SCOPE = ("GPD", "PointNetGPD")


def build_model(cfg):
    kind = cfg.MODEL.TYPE
    if kind in _BUILDERS:
        return _BUILDERS[kind](cfg)
    if kind in _OUT_OF_SCOPE:
        raise ValueError("model %r is one of the reference's baselines (no PointNet++ operators); it is not part of "
                         "this B200 build — use the reference implementation for it" % kind)
    raise ValueError("Unknown model: {}.".format(kind))
