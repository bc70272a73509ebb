"""GPU suite: the fp64 instantiation of the seven pn2_ext operators (the reference dispatches float AND double,
e.g. sampling_kernel.cu:21) — reference CUDA kernels (oracle/_ref, when built) vs the C oracle vs the sm_100a
kernels, bit-exact for indices, distances, gathers and interpolation; atomics-ordered sums within 1e-12."""
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
        return None
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
    x = {"uniform": lambda: inputs.uniform_cloud(B, N, seed), "lattice": lambda: inputs.lattice_cloud(B, N, seed, side=8),
         "dup": lambda: inputs.duplicated_cloud(B, N, seed), "identical": lambda: inputs.identical_cloud(B, N),
         "scene": lambda: inputs.tabletop_batch(B, 1000 + seed, N)}[kind]().double()
    if kind == "uniform":  # use the extra mantissa bits: values that are NOT representable in fp32
        x = x + torch.from_numpy(np.random.RandomState(seed + 1).rand(B, 3, N)) * 1e-9
    return x


@pytest.mark.parametrize("kind,B,N,M", [("uniform", 2, 5120, 1024), ("lattice", 3, 2000, 600), ("lattice", 2, 40, 40),
                                        ("dup", 2, 3000, 1500), ("identical", 2, 600, 50), ("scene", 1, 25600, 700),
                                        ("uniform", 2, 13, 9), ("lattice", 1, 30000, 300)])
def test_fps_f64(ref, ext, ora, kind, B, N, M):
    pts = _gen(kind, B, N, seed=7)
    o = ora.farthest_point_sample(pts, M)
    assert torch.equal(ora.farthest_point_sample(pts, M, keyed=True), o)
    if ref is not None:
        assert torch.equal(ref.farthest_point_sample(pts.cuda(), M).cpu(), o), "C oracle differs from the reference"
    got = ext.farthest_point_sample(pts.cuda(), M)
    assert got.dtype == torch.int64 and torch.equal(got.cpu(), o)


@pytest.mark.parametrize("kind,B,N,M,rad,K", [("uniform", 2, 5120, 512, 0.08, 64), ("scene", 1, 25600, 640, 0.02, 64),
                                              ("lattice", 2, 2000, 500, 0.125, 32), ("dup", 2, 3000, 400, 0.05, 32),
                                              ("uniform", 2, 50, 3, 1e-4, 8), ("identical", 2, 300, 20, 0.1, 16)])
def test_ball_query_f64(ref, ext, ora, kind, B, N, M, rad, K):
    pts = _gen(kind, B, N, seed=8)
    ctr = ora.gather_points(pts, ora.farthest_point_sample(pts, M))
    o_idx, o_cnt = ora.ball_query(pts, ctr, rad, K)
    if ref is not None:
        r_idx, r_cnt = [t.cpu() for t in ref.ball_query(pts.cuda(), ctr.cuda(), rad, K)]
        assert torch.equal(o_idx, r_idx) and torch.equal(o_cnt, r_cnt), "C oracle differs from the reference"
    g_idx, g_cnt = [t.cpu() for t in ext.ball_query(pts.cuda(), ctr.cuda(), rad, K)]
    assert torch.equal(g_idx, o_idx) and torch.equal(g_cnt, o_cnt)


@pytest.mark.parametrize("kind,B,Nq,Nk", [("uniform", 2, 5120, 1024), ("lattice", 2, 3000, 300), ("dup", 2, 2000, 900),
                                          ("uniform", 3, 101, 3), ("uniform", 1, 300, 2500)])
def test_point_search_f64(ref, ext, ora, kind, B, Nq, Nk):
    q = _gen(kind, B, Nq, seed=9)
    k = _gen(kind, B, Nk, seed=10)
    o_idx, o_d = ora.point_search(q, k, 3)
    if ref is not None:
        r_idx, r_d = [t.cpu() for t in ref.point_search(q.cuda(), k.cuda(), 3)]
        assert torch.equal(o_idx, r_idx) and torch.equal(o_d, r_d), "C oracle differs from the reference"
    g_idx, g_d = [t.cpu() for t in ext.point_search(q.cuda(), k.cuda(), 3)]
    assert g_d.dtype == torch.float64
    assert torch.equal(g_idx, o_idx) and torch.equal(g_d, o_d)


def test_group_and_interpolate_f64(ref, ext, ora):
    rs = np.random.RandomState(1)
    B, C, N, M, K = 2, 19, 900, 50, 16
    x = torch.from_numpy(rs.randn(B, C, N))
    idx = torch.from_numpy(rs.randint(0, N, size=(B, M, K)).astype(np.int64))
    o = ora.group_points_forward(x, idx)
    if ref is not None:
        assert torch.equal(ref.group_points_forward(x.cuda(), idx.cuda()).cpu(), o)
    assert torch.equal(ext.group_points_forward(x.cuda(), idx.cuda()).cpu(), o)
    g = torch.from_numpy(rs.randn(B, C, M, K))
    o = ora.group_points_backward(g, idx, N)
    if ref is not None:
        np.testing.assert_allclose(ref.group_points_backward(g.cuda(), idx.cuda(), N).cpu().numpy(), o.numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(ext.group_points_backward(g.cuda(), idx.cuda(), N).cpu().numpy(), o.numpy(), rtol=1e-12, atol=1e-12)
    idx3 = torch.from_numpy(rs.randint(0, N, size=(B, M, 3)).astype(np.int64))
    w = torch.from_numpy(rs.rand(B, M, 3))
    o = ora.interpolate_forward(x, idx3, w)
    if ref is not None:
        assert torch.equal(ref.interpolate_forward(x.cuda(), idx3.cuda(), w.cuda()).cpu(), o)
    assert torch.equal(ext.interpolate_forward(x.cuda(), idx3.cuda(), w.cuda()).cpu(), o)
    g2 = torch.from_numpy(rs.randn(B, C, M))
    o = ora.interpolate_backward(g2, idx3, w, N)
    if ref is not None:
        np.testing.assert_allclose(ref.interpolate_backward(g2.cuda(), idx3.cuda(), w.cuda(), N).cpu().numpy(), o.numpy(),
                                   rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(ext.interpolate_backward(g2.cuda(), idx3.cuda(), w.cuda(), N).cpu().numpy(), o.numpy(),
                               rtol=1e-12, atol=1e-12)


def test_mixed_dtypes_raise(ext):
    p = torch.rand(1, 3, 64, device="cuda", dtype=torch.float64)
    with pytest.raises(RuntimeError):
        ext.ball_query(p, p.float(), 0.1, 4)
    with pytest.raises(RuntimeError):
        ext.farthest_point_sample(p.half(), 4)


def test_autograd_wrappers_f64(ext):
    """functions.py:28-174 wrappers run in double end to end (gradcheck-style comparison with torch ops)."""
    from s4g_release_b200.network_models.models.pointnet2_utils import functions as F
    rs = np.random.RandomState(3)
    B, C, N, M, K = 2, 5, 60, 12, 4
    x = torch.from_numpy(rs.randn(B, C, N)).cuda().requires_grad_(True)
    idx = torch.from_numpy(rs.randint(0, N, size=(B, M, K)).astype(np.int64)).cuda()
    out = F.group_points(x, idx)
    ref = torch.gather(x.unsqueeze(2).expand(-1, -1, M, -1), 3, idx.unsqueeze(1).expand(-1, C, -1, -1))
    assert out.dtype == torch.float64 and torch.equal(out, ref)
    gy = torch.from_numpy(rs.randn(B, C, M, K)).cuda()
    (gx,) = torch.autograd.grad(out, x, gy)
    (gr,) = torch.autograd.grad(ref, x, gy)
    np.testing.assert_allclose(gx.cpu().numpy(), gr.cpu().numpy(), rtol=1e-12, atol=1e-12)
