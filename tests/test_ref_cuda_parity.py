"""GPU suite: the REFERENCE's own CUDA kernels (compiled unmodified for sm_100a into oracle/_ref by
oracle/build_ref.py) side by side with (a) the C oracle, which pins the oracle, and (b) the sm_100a
kernels of this repo.  Bit-exact for FPS / ball query / 3-NN / gathers / interpolation."""
import os

import numpy as np
import pytest
import torch

from tests import inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    from oracle import build_ref
    if not os.path.exists(build_ref.so_path()):
        pytest.skip("oracle/_ref/ref_pn2_ext.so not built (needs /root/reference in the build container)")
    return build_ref.load()


@pytest.fixture(scope="module")
def ext():
    from s4g_release_b200.network_models.models.pointnet2_utils import pn2_ext
    return pn2_ext


@pytest.fixture(scope="module")
def ora():
    from oracle import pn2_ext_cpu
    return pn2_ext_cpu


def _gen(kind, B, N, seed):
    return {"uniform": lambda: inputs.uniform_cloud(B, N, seed), "lattice": lambda: inputs.lattice_cloud(B, N, seed, side=8),
            "dup": lambda: inputs.duplicated_cloud(B, N, seed), "identical": lambda: inputs.identical_cloud(B, N),
            "scene": lambda: inputs.tabletop_batch(B, 1000 + seed, N)}[kind]()


@pytest.mark.parametrize("kind,B,N,M", [("uniform", 2, 5120, 1024), ("lattice", 3, 2000, 600), ("lattice", 2, 200, 150),
                                        ("lattice", 2, 40, 40), ("dup", 2, 3000, 1500), ("identical", 2, 600, 50),
                                        ("scene", 2, 25600, 5120), ("uniform", 2, 13, 9), ("lattice", 1, 30000, 1500)])
def test_fps_three_way(ref, ext, ora, kind, B, N, M):
    pts = _gen(kind, B, N, seed=7)
    r = ref.farthest_point_sample(pts.cuda(), M).cpu()
    assert torch.equal(ora.farthest_point_sample(pts, M), r), "C oracle differs from the reference CUDA kernel"
    assert torch.equal(ext.farthest_point_sample(pts.cuda(), M).cpu(), r), "sm_100a kernel differs from the reference"


@pytest.mark.parametrize("kind,B,N,M,rad,K", [("uniform", 2, 5120, 1024, 0.08, 64), ("scene", 1, 25600, 5120, 0.02, 64),
                                              ("lattice", 2, 2000, 500, 0.125, 32), ("dup", 2, 3000, 400, 0.05, 32),
                                              ("uniform", 2, 50, 3, 1e-4, 8), ("identical", 2, 300, 20, 0.1, 16)])
def test_ball_query_three_way(ref, ext, ora, kind, B, N, M, rad, K):
    pts = _gen(kind, B, N, seed=8)
    sel = ora.farthest_point_sample(pts, M)
    ctr = ora.gather_points(pts, sel)
    r_idx, r_cnt = [t.cpu() for t in ref.ball_query(pts.cuda(), ctr.cuda(), rad, K)]
    o_idx, o_cnt = ora.ball_query(pts, ctr, rad, K)
    assert torch.equal(o_idx, r_idx) and torch.equal(o_cnt, r_cnt), "C oracle differs from the reference CUDA kernel"
    g_idx, g_cnt = [t.cpu() for t in ext.ball_query(pts.cuda(), ctr.cuda(), rad, K)]
    assert torch.equal(g_idx, r_idx) and torch.equal(g_cnt, r_cnt), "sm_100a kernel differs from the reference"


@pytest.mark.parametrize("kind,B,Nq,Nk", [("uniform", 2, 5120, 1024), ("scene", 1, 25600, 5120), ("lattice", 2, 3000, 300),
                                          ("dup", 2, 2000, 900), ("uniform", 3, 101, 3)])
def test_point_search_three_way(ref, ext, ora, kind, B, Nq, Nk):
    q = _gen(kind, B, Nq, seed=9)
    k = ora.gather_points(q, ora.farthest_point_sample(q, Nk)) if kind == "scene" else _gen(kind, B, Nk, seed=10)
    r_idx, r_d = [t.cpu() for t in ref.point_search(q.cuda(), k.cuda(), 3)]
    o_idx, o_d = ora.point_search(q, k, 3)
    assert torch.equal(o_idx, r_idx) and torch.equal(o_d, r_d), "C oracle differs from the reference CUDA kernel"
    g_idx, g_d = [t.cpu() for t in ext.point_search(q.cuda(), k.cuda(), 3)]
    assert torch.equal(g_idx, r_idx) and torch.equal(g_d, r_d), "sm_100a kernel differs from the reference"


def test_group_and_interpolate_three_way(ref, ext, ora):
    rs = np.random.RandomState(1)
    B, C, N, M, K = 2, 67, 900, 50, 16
    x = torch.from_numpy(rs.randn(B, C, N).astype(np.float32))
    idx = torch.from_numpy(rs.randint(0, N, size=(B, M, K)).astype(np.int64))
    r = ref.group_points_forward(x.cuda(), idx.cuda()).cpu()
    assert torch.equal(ora.group_points_forward(x, idx), r)
    assert torch.equal(ext.group_points_forward(x.cuda(), idx.cuda()).cpu(), r)
    g = torch.from_numpy(rs.randn(B, C, M, K).astype(np.float32))
    r = ref.group_points_backward(g.cuda(), idx.cuda(), N).cpu()
    np.testing.assert_allclose(ora.group_points_backward(g, idx, N).numpy(), r.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(ext.group_points_backward(g.cuda(), idx.cuda(), N).cpu().numpy(), r.numpy(), rtol=1e-4, atol=1e-5)
    idx3 = torch.from_numpy(rs.randint(0, N, size=(B, M, 3)).astype(np.int64))
    w = torch.from_numpy(rs.rand(B, M, 3).astype(np.float32))
    r = ref.interpolate_forward(x.cuda(), idx3.cuda(), w.cuda()).cpu()
    assert torch.equal(ora.interpolate_forward(x, idx3, w), r)
    assert torch.equal(ext.interpolate_forward(x.cuda(), idx3.cuda(), w.cuda()).cpu(), r)
    g2 = torch.from_numpy(rs.randn(B, C, M).astype(np.float32))
    r = ref.interpolate_backward(g2.cuda(), idx3.cuda(), w.cuda(), N).cpu()
    np.testing.assert_allclose(ora.interpolate_backward(g2, idx3, w, N).numpy(), r.numpy(), rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(ext.interpolate_backward(g2.cuda(), idx3.cuda(), w.cuda(), N).cpu().numpy(), r.numpy(),
                               rtol=1e-4, atol=1e-5)
