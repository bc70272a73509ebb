"""Device-side cloud pre-processing — host mirror of GraspDetector._pre_processing / sample_single_cloud
(grasp_detector.py:82-105) and of the geometric part of CloudPreProcessor (cloud_processor/cloud_processor.py:13-42)
on the kernels of csrc/preprocess.cu and the ball-query kernel.  CUDA tensors only (no CPU path).

What the reference observably does: ``voxelize()`` / ``remove_outliers()`` drop the clouds open3d returns (no-ops), then
the cloud is mapped by _REAL2TRAIN and randomly sub-sampled to NUM_INPUT points.  ``GraspPreProcessor.pre_processing``
reproduces exactly that (``apply_filters=False``, the default); ``apply_filters=True`` runs the two filters the code
intends, with the definitions stated in csrc/preprocess.cu.
"""
import ctypes

import numpy as np
import torch

from ._lib import check, lib, ptr, stream_ptr

REAL2TRAIN = np.array([[0, 1, 0, 0], [1, 0, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1]], dtype=np.float32)  # grasp_detector.py:26
VOXEL_SIZE, NUM_POINTS_THRESHOLD, RADIUS_THRESHOLD = 0.005, 32, 0.02  # configs/processing_config.py:20-22
NUM_INPUT = 25600  # cfg.MODEL.PN2.NUM_INPUT (configs/curvature_model.yaml)


def _need_cuda(t):
    if not (torch.is_tensor(t) and t.is_cuda):
        raise RuntimeError("cloud pre-processing needs CUDA tensors (there is no CPU path)")


def transform_select(clouds, index=None, matrix=REAL2TRAIN):
    """clouds (B,3,n) fp32, index (B,m) int64 or None -> (B,3,m): transform_numpy_points (utils/math_utils.py:20-24)
    followed by ``points[:, random_index]`` (grasp_detector.py:91) in one pass."""
    _need_cuda(clouds)
    x = clouds.float().contiguous()
    B, _, n = x.shape
    m = n if index is None else index.shape[1]
    if index is not None:
        index = index.to(device=x.device, dtype=torch.int64).contiguous()
        if index.numel() and (int(index.min()) < 0 or int(index.max()) >= n):
            raise RuntimeError("index out of range")
    mat = np.ascontiguousarray(np.asarray(matrix, dtype=np.float32).reshape(4, 4))
    out = torch.empty((B, 3, m), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(lib.s4g_cloud_transform_select_f32(ptr(x), B, n, ptr(index) if index is not None else None, m,
                                                 mat.ctypes.data_as(ctypes.c_void_p), ptr(out), stream_ptr(x.device)),
              "cloud_transform_select")
    return out


class CloudPreProcessor:
    """cloud_processor.CloudPreProcessor on a (3, n) CUDA tensor.  Like open3d, ``voxelize`` and ``remove_outliers``
    RETURN the filtered cloud and leave ``self.points`` alone (which is why they are no-ops in the reference)."""

    def __init__(self, cloud_3n):
        _need_cuda(cloud_3n)
        self.points = cloud_3n.float().contiguous()

    def filter_work_space(self, workspace):
        """cloud_processor.py:13-29: strict inequalities on all six bounds; modifies the cloud, returns the mask."""
        p = self.points
        lo = torch.tensor([workspace[0], workspace[2], workspace[4]], device=p.device).view(3, 1)
        hi = torch.tensor([workspace[1], workspace[3], workspace[5]], device=p.device).view(3, 1)
        valid = ((p > lo) & (p < hi)).all(dim=0)
        self.points = p[:, valid].contiguous()
        return valid

    def voxelize(self, voxel_size=VOXEL_SIZE):
        """One point per occupied voxel: the mean of its points, voxels in ascending (z, y, x) cell order."""
        p = self.points
        n = p.shape[1]
        origin = (p.min(dim=1).values - 0.5 * voxel_size).cpu().numpy().astype(np.float32)  # open3d's min_bound - voxel/2
        extent = p.max(dim=1).values.cpu().numpy() - origin
        dims = np.maximum(np.floor(extent / voxel_size).astype(np.int32) + 1, 1).astype(np.int32)
        key = torch.empty(n, dtype=torch.int64, device=p.device)
        with torch.cuda.device(p.device):
            check(lib.s4g_voxel_keys_f32(ptr(p), n, origin.ctypes.data_as(ctypes.c_void_p), float(voxel_size),
                                         dims.ctypes.data_as(ctypes.c_void_p), ptr(key), stream_ptr(p.device)), "voxel_keys")
            skey, order = torch.sort(key, stable=True)
            start = torch.ones(n, dtype=torch.int32, device=p.device)
            start[1:] = (skey[1:] != skey[:-1]).int()
            rank = (torch.cumsum(start, 0) - start).int().contiguous()
            n_vox = int(start.sum())
            out = torch.empty((3, n_vox), dtype=torch.float32, device=p.device)
            check(lib.s4g_voxel_means_f32(ptr(p), n, ptr(skey), ptr(order), ptr(rank), ptr(out), n_vox,
                                          stream_ptr(p.device)), "voxel_means")
        return out

    def remove_outliers(self, nb_points=NUM_POINTS_THRESHOLD, radius=RADIUS_THRESHOLD):
        """Keep the points with MORE than nb_points points (itself included) within ``radius``.  Returns (cloud, mask)."""
        from .network_models.models.pointnet2_utils import pn2_ext
        p = self.points.unsqueeze(0)
        _, count = pn2_ext.ball_query(p, p, radius, nb_points + 1)
        mask = count[0] > nb_points
        return self.points[:, mask].contiguous(), mask


class GraspPreProcessor:
    """GraspDetector._pre_processing + sample_single_cloud (grasp_detector.py:82-105)."""

    def __init__(self, num_input=NUM_INPUT, apply_filters=False):
        self.num_input, self.apply_filters = int(num_input), bool(apply_filters)

    def sample_index(self, n, rng=np.random):
        """sample_single_cloud's index draw (grasp_detector.py:86-89): without replacement when the cloud is larger."""
        return rng.choice(np.arange(n), self.num_input, replace=not (n > self.num_input))

    def pre_processing(self, cloud_array, rng=np.random, random_index=None):
        """cloud (3, n) -> points (3, num_input) fp32 CUDA tensor in the training frame."""
        cloud = cloud_array if torch.is_tensor(cloud_array) else torch.as_tensor(np.asarray(cloud_array, dtype=np.float32))
        cloud = cloud.cuda() if not cloud.is_cuda else cloud
        if self.apply_filters:
            pre = CloudPreProcessor(cloud)
            pre.points = pre.voxelize()
            pre.points, _ = pre.remove_outliers()
            cloud = pre.points
        if random_index is None:
            random_index = self.sample_index(cloud.shape[1], rng)
        idx = torch.as_tensor(np.asarray(random_index), dtype=torch.int64).view(1, -1)
        return transform_select(cloud.unsqueeze(0), idx)[0]

    def pre_processing_batch(self, clouds, random_index):
        """clouds (B,3,n), random_index (B, num_input) -> (B,3,num_input): a whole batch in one launch."""
        return transform_select(clouds, torch.as_tensor(random_index))
