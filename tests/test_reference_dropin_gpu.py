"""GPU: the Level-1 drop-in test.  The UNMODIFIED reference python — its own
PointNet2_tcls.PointNet2, pointnet2_utils/modules.py + functions.py, nn_utils/* staged byte for byte under
baseline/_ref by baseline/stage_ref.py — runs on
  (i)  the reference's own CUDA extension (oracle/_ref/ref_pn2_ext.so: its .cu files compiled unmodified for sm_100a),
  (ii) this repo's ``pn2_ext`` (the C-ABI library behind the reference's seven names),
and must give identical indices and (the rest being the same torch code) identical outputs; the product's module path
agrees to fp32 tolerance."""
import os

import numpy as np
import pytest
import torch

from tests.inputs import TINY_CONFIG

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def staged():
    from baseline import stage_ref
    from oracle import build_ref
    if not stage_ref.available():
        pytest.skip("baseline/_ref not staged (python baseline/stage_ref.py in the build container)")
    if not os.path.exists(build_ref.so_path()):
        pytest.skip("oracle/_ref/ref_pn2_ext.so not built")
    return stage_ref, build_ref.load()


class _Recorder:
    """pn2_ext stand-in that forwards to `impl` and records the index outputs."""

    def __init__(self, impl):
        self.impl, self.log = impl, []
        for fn in ("group_points_forward", "group_points_backward", "interpolate_forward", "interpolate_backward"):
            setattr(self, fn, getattr(impl, fn))

    def farthest_point_sample(self, *a):
        out = self.impl.farthest_point_sample(*a)
        self.log.append(("fps", out.clone()))
        return out

    def ball_query(self, *a):
        out = self.impl.ball_query(*a)
        self.log.append(("ball", out[0].clone()))
        return out

    def point_search(self, *a):
        out = self.impl.point_search(*a)
        self.log.append(("nn", out[0].clone()))
        return out


@pytest.mark.parametrize("config", ["tiny", "pn2_cls"])
def test_reference_model_runs_unmodified_on_our_ops(staged, golden_tiny, cloud_2638, config):
    stage_ref, ref_ext = staged
    from s4g_release_b200.network_models.models.pointnet2_utils import pn2_ext as our_ext
    from s4g_release_b200.network_models.models.PointNet2_tcls import PN2_CLS_CONFIG, PointNet2 as OurPointNet2
    from tests.golden.make_golden import seed_reference_weights
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = TINY_CONFIG if config == "tiny" else PN2_CLS_CONFIG
    pts = torch.from_numpy(golden_tiny["points"] if config == "tiny" else cloud_2638[None]).cuda()
    outs, logs = [], []
    for impl in (ref_ext, our_ext):
        rec = _Recorder(impl)
        RefPointNet2 = stage_ref.import_reference_model(rec)
        torch.manual_seed(0)
        model = seed_reference_weights(RefPointNet2(**cfg)).cuda().eval()
        with torch.no_grad():
            outs.append(model({"scene_points": pts}))
        torch.cuda.synchronize()
        logs.append(rec.log)
    assert [k for k, _ in logs[0]] == [k for k, _ in logs[1]] and len(logs[0]) == 9  # 3 x (fps, ball) + 3 x 3-NN
    for (kind, a), (_, b) in zip(*logs):
        assert a.dtype == b.dtype == torch.int64 and torch.equal(a, b), "%s indices differ from the reference's CUDA op" % kind
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), "reference model output %s differs between the two op sets" % k
    # the product's module path (same parameters) against the unmodified reference model
    torch.manual_seed(0)
    ours = seed_reference_weights(OurPointNet2(**cfg)).cuda().eval()
    with torch.no_grad():
        mine = ours({"scene_points": pts}, fused=False)
    for k in outs[0]:
        err = (mine[k] - outs[0][k]).abs().max().item()
        assert err <= 2e-3 * max(1.0, outs[0][k].abs().max().item()), (k, err)


def test_reference_autograd_wrappers_backward_through_our_ops(staged):
    """the reference's functions.py autograd wrappers (GroupPoints / FeatureInterpolate backward) on our ops: gradients
    equal those on the reference's CUDA ops up to atomicAdd order"""
    stage_ref, ref_ext = staged
    from s4g_release_b200.network_models.models.pointnet2_utils import pn2_ext as our_ext
    import importlib
    g = torch.Generator().manual_seed(2)
    xyz = torch.rand(2, 3, 512, generator=g).cuda()
    feat = torch.randn(2, 16, 512, generator=g).cuda()
    grads = []
    for impl in (ref_ext, our_ext):
        stage_ref.import_reference_model(impl)
        F = importlib.import_module("grasp_proposal.network_models.models.pointnet2_utils.functions")
        f = feat.clone().requires_grad_(True)
        idx = F.farthest_point_sample(xyz, 64)
        ctr = F.gather_points(xyz, idx)
        nbr, _ = F.ball_query(xyz, ctr, 0.2, 8)
        grouped = F.group_points(f, nbr)
        nn_idx, d2 = F.search_nn_distance(xyz, ctr, 3)
        w = 1.0 / torch.clamp(d2, min=1e-10)
        w = w / w.sum(2, keepdim=True)
        interp = F.feature_interpolate(grouped.max(3)[0], nn_idx, w)
        (interp ** 2).sum().backward()
        grads.append(f.grad.clone())
    assert torch.allclose(grads[0], grads[1], rtol=1e-4, atol=1e-5)
