"""GPU parity of the device-side pre-processing (csrc/preprocess.cu, s4g_release_b200/preprocess.py) against the CPU
restatement of GraspDetector._pre_processing (oracle/model_cpu.py): bit-exact for the reference-observable path
(axis change + sub-sample), exact cell membership and 1e-6 means for the voxel filter, exact mask for the outlier filter."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cloud(n, seed):
    rs = np.random.RandomState(seed)
    return (rs.rand(3, n).astype(np.float32) * np.array([[0.8], [0.7], [0.3]], dtype=np.float32)
            + np.array([[-0.4], [-0.35], [0.75]], dtype=np.float32))


@pytest.mark.parametrize("n,m", [(48902, 25600), (9000, 25600), (25600, 25600), (10, 64)])
def test_pre_processing_matches_reference_behaviour(n, m):
    from oracle import model_cpu as ora
    from s4g_release_b200.preprocess import GraspPreProcessor
    cloud = _cloud(n, n)
    pre = GraspPreProcessor(num_input=m)
    idx = pre.sample_index(n, np.random.RandomState(3))
    assert len(idx) == m and (len(set(idx.tolist())) == m) == (n > m)  # without replacement only when n > m (:86-89)
    got = pre.pre_processing(torch.from_numpy(cloud).cuda(), random_index=idx)
    want = ora.pre_processing(cloud, idx)
    assert got.dtype == torch.float32 and tuple(got.shape) == (3, m)
    assert np.array_equal(got.cpu().numpy(), want)


def test_pre_processing_batch_and_fixture(cloud_2638):
    from oracle import model_cpu as ora
    from s4g_release_b200.preprocess import GraspPreProcessor
    rs = np.random.RandomState(0)
    base = np.ascontiguousarray(np.asarray(cloud_2638, dtype=np.float32).reshape(3, -1))
    clouds = np.stack([base + np.float32(0.001 * b) for b in range(3)])
    idx = np.stack([rs.choice(base.shape[1], 2048, replace=False) for _ in range(3)])
    got = GraspPreProcessor(num_input=2048).pre_processing_batch(torch.from_numpy(clouds).cuda(), idx).cpu().numpy()
    for b in range(3):
        assert np.array_equal(got[b], ora.pre_processing(clouds[b], idx[b]))


def test_index_out_of_range_raises():
    from s4g_release_b200.preprocess import transform_select
    c = torch.rand(1, 3, 100, device="cuda")
    with pytest.raises(RuntimeError):
        transform_select(c, torch.tensor([[0, 100]]))
    with pytest.raises(RuntimeError):
        transform_select(c.cpu())


@pytest.mark.parametrize("n,voxel", [(20000, 0.005), (3000, 0.02), (500, 0.5)])
def test_voxel_filter(n, voxel):
    from oracle import model_cpu as ora
    from s4g_release_b200.preprocess import CloudPreProcessor
    cloud = _cloud(n, 7)
    pre = CloudPreProcessor(torch.from_numpy(cloud).cuda())
    got = pre.voxelize(voxel).cpu().numpy()
    want = ora.voxel_down_sample(cloud, voxel)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, atol=1e-6, rtol=0)
    assert torch.equal(pre.points.cpu(), torch.from_numpy(cloud))  # like open3d: returns, does not modify


def test_outlier_filter_and_workspace():
    from oracle import model_cpu as ora
    from s4g_release_b200.preprocess import CloudPreProcessor
    cloud = _cloud(6000, 9)
    cloud[:, :50] += 5.0  # isolated points
    pre = CloudPreProcessor(torch.from_numpy(cloud).cuda())
    kept, mask = pre.remove_outliers(nb_points=4, radius=0.03)
    want = ora.radius_outlier_mask(cloud, 4, 0.03)
    assert np.array_equal(mask.cpu().numpy(), want) and not want[:50].any() and want.any()
    assert kept.shape[1] == int(want.sum())
    ws = [-0.2, 0.2, -0.1, 0.3, 0.8, 1.0]
    valid = pre.filter_work_space(ws).cpu().numpy()
    p = cloud
    ref = (p[0] > ws[0]) & (p[0] < ws[1]) & (p[1] > ws[2]) & (p[1] < ws[3]) & (p[2] > ws[4]) & (p[2] < ws[5])
    assert np.array_equal(valid, ref) and pre.points.shape[1] == int(ref.sum())


def test_intended_filters_pipeline_runs():
    from s4g_release_b200.preprocess import GraspPreProcessor
    cloud = _cloud(30000, 11)
    cloud[2] = 0.75 + 0.002 * cloud[2]  # a table-like slab: dense enough to survive the 32-points-in-2-cm filter
    pts = GraspPreProcessor(num_input=4096, apply_filters=True).pre_processing(torch.from_numpy(cloud).cuda(),
                                                                                rng=np.random.RandomState(0))
    assert tuple(pts.shape) == (3, 4096) and torch.isfinite(pts).all()


def test_against_reference_golden():
    import os
    from s4g_release_b200.preprocess import GraspPreProcessor
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "preprocess_ref.npz"))
    for name in ("large", "small"):
        cloud, index = g[name + "/cloud"], g[name + "/index"]
        pre = GraspPreProcessor(num_input=len(index))
        assert np.array_equal(pre.sample_index(cloud.shape[1], np.random.RandomState(17)), index)
        got = pre.pre_processing(torch.from_numpy(cloud).cuda(), random_index=index)
        assert np.array_equal(got.cpu().numpy(), g[name + "/points_f32"])
