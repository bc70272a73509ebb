"""EDGEPN2D — PN2 with EdgeConv set-abstraction levels (reference: network_models/models/EdgePointNet2Down.py:9-33).

The reference's class is the PN2 sibling with ``_SA_MODULE = EdgeSAModule`` and the plain PointnetFPModule decoder; the
builder reads the ``MODEL.EDGEPN2D`` node.  Edge set abstraction concatenates centroid features to every neighbour, which
the fused engine has no plan for: this model always runs on the module path (sm_100a operators + torch autograd)."""
from . import PointNet2 as _pn2
from .pointnet2_utils.modules import EdgeSAModule, PointnetFPModule

PointNet2Loss, PointNet2Metric = _pn2.PointNet2Loss, _pn2.PointNet2Metric


class EdgePointNet2Down(_pn2.PointNet2):
    _SA_MODULE = EdgeSAModule
    _FP_MODULE = PointnetFPModule

    def fusable(self):
        return False


def build_edgepointnet2down(cfg):
    node = cfg.MODEL.EDGEPN2D
    net = EdgePointNet2Down(score_classes=cfg.DATA.SCORE_CLASSES, num_centroids=node.NUM_CENTROIDS, radius=node.RADIUS,
                            num_neighbours=node.NUM_NEIGHBOURS, sa_channels=node.SA_CHANNELS, fp_channels=node.FP_CHANNELS,
                            num_fp_neighbours=node.NUM_FP_NEIGHBOURS, seg_channels=node.SEG_CHANNELS,
                            dropout_prob=node.DROPOUT_PROB)
    return net, PointNet2Loss(label_smoothing=node.LABEL_SMOOTHING, neg_weight=node.NEG_WEIGHT), PointNet2Metric()
