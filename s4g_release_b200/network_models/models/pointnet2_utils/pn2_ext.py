"""Drop-in replacement of the reference's compiled ``pn2_ext`` module.

Same seven names, argument order, channel-first shapes, int64 indices and RuntimeError behaviour as
inference/grasp_proposal/network_models/models/pointnet2_utils/csrc/main.cpp:7-13, implemented by
the hand-written sm_100a kernels of ``libs4g_b200.so`` (include/s4g_b200.h).  Unlike the reference it
launches on torch's CURRENT stream and under a device guard; inputs may be non-contiguous (the
reference transposes + ``.contiguous()`` itself, e.g. sampling_kernel.cu:141) and are never modified.

To use it under the unmodified reference package::

    import sys
    from s4g_release_b200.network_models.models.pointnet2_utils import pn2_ext
    sys.modules["grasp_proposal.network_models.models.pointnet2_utils.pn2_ext"] = pn2_ext
"""
import ctypes

import torch

from ...._lib import check, lib, ptr, stream_ptr


def _cuda_f32(t, name, like=None):
    """float32 or float64 CUDA tensor (the reference dispatches both, e.g. sampling_kernel.cu:21); every floating
    argument of one call must have the dtype of the first (`like`)."""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)  # CHECK_CUDA
    if t.dtype not in (torch.float32, torch.float64):
        raise RuntimeError("%s must be float32 or float64 (got %s)" % (name, t.dtype))
    if like is not None and t.dtype != like.dtype:
        raise RuntimeError("%s: expected %s, got %s" % (name, like.dtype, t.dtype))
    return t.contiguous()


def _fn(name, t):
    return getattr(lib, name + ("_f64" if t.dtype == torch.float64 else "_f32"))


def _cuda_i64(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if t.dtype != torch.int64:
        raise RuntimeError("%s must be int64 (got %s)" % (name, t.dtype))
    return t.contiguous()


def _require(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def farthest_point_sample(points, num_centroids):
    """(B,3,N) fp32 -> (B,M) int64.  sampling.h:7-9 / sampling_kernel.cu:128-172."""
    points = _cuda_f32(points, "points")
    _require(points.dim() == 3 and points.size(1) == 3, "points.size(1) != 3")
    B, _, N = points.shape
    M = int(num_centroids)
    _require(M > 0, "num_centroids <= 0")
    _require(N >= M, "num_points < num_centroids")
    with torch.cuda.device(points.device):
        index = torch.empty((B, M), dtype=torch.int64, device=points.device)
        if points.dtype == torch.float64:
            ws_bytes = int(lib.s4g_farthest_point_sample_f64_workspace(B, N))
            ws = torch.empty(max(ws_bytes, 8), dtype=torch.uint8, device=points.device)
            check(lib.s4g_farthest_point_sample_f64(ptr(points), B, N, M, ptr(index), ptr(ws), ws_bytes,
                                                    stream_ptr(points.device)), "farthest_point_sample")
        else:
            check(lib.s4g_farthest_point_sample_f32(ptr(points), B, N, M, ptr(index), stream_ptr(points.device)),
                  "farthest_point_sample")
    return index


def ball_query(points, centroids, radius, num_neighbours):
    """(B,3,N),(B,3,M) -> [index (B,M,K) int64, count (B,M) int64].  ball_query.h:7-11."""
    points = _cuda_f32(points, "points")
    centroids = _cuda_f32(centroids, "centroids", points)
    _require(points.dim() == 3 and points.size(1) == 3, "points.size(1) != 3")
    _require(centroids.dim() == 3 and centroids.size(1) == 3, "centroids.size(1) != 3")
    _require(centroids.size(0) == points.size(0), "centroids.size(0) != batch_size")
    B, _, N = points.shape
    M = centroids.size(2)
    K = int(num_neighbours)
    with torch.cuda.device(points.device):
        index = torch.empty((B, M, K), dtype=torch.int64, device=points.device)
        count = torch.empty((B, M), dtype=torch.int64, device=points.device)
        radius = ctypes.c_float(radius).value  # the reference's argument is a C float (ball_query.h:10)
        check(_fn("s4g_ball_query", points)(ptr(points), ptr(centroids), B, N, M, radius, K, ptr(index), ptr(count),
                                            stream_ptr(points.device)), "ball_query")
    return [index, count]


def group_points_forward(input, index):
    """(B,C,N),(B,M,K) -> (B,C,M,K).  grouping.h:7-9 / grouping_kernel.cu:32-54."""
    input = _cuda_f32(input, "input")
    index = _cuda_i64(index, "index")
    _require(input.dim() == 3, "input.dim() != 3")
    _require(index.dim() == 3, "index.dim() != 3")
    _require(index.size(0) == input.size(0), "index.size(0) != batch_size")
    B, C, N = input.shape
    _, M, K = index.shape
    with torch.cuda.device(input.device):
        out = torch.empty((B, C, M, K), dtype=input.dtype, device=input.device)
        check(_fn("s4g_group_points_forward", input)(ptr(input), ptr(index), B, C, N, M, K, ptr(out),
                                               stream_ptr(input.device)), "group_points_forward")
    return out


def gather_points(points, index):
    """(B,C,N) fp32, (B,M) int64 -> (B,C,M): out[b,c,m] = points[b,c,index[b,m]].  Not one of the reference's seven
    extension functions — its ``functions.gather_points`` (functions.py:10-25) is a torch.gather; this is the same
    result from the library's kernel for the no-autograd path."""
    points = _cuda_f32(points, "points")
    index = _cuda_i64(index, "index")
    _require(points.dtype == torch.float32, "gather_points: float32 only")
    _require(points.dim() == 3 and index.dim() == 2 and index.size(0) == points.size(0), "gather_points: bad shapes")
    B, C, N = points.shape
    M = index.size(1)
    with torch.cuda.device(points.device):
        out = torch.empty((B, C, M), dtype=points.dtype, device=points.device)
        check(lib.s4g_gather_points_f32(ptr(points), ptr(index), B, C, N, M, ptr(out), stream_ptr(points.device)),
              "gather_points")
    return out


def group_points_backward(grad_output, index, num_points):
    """(B,C,M,K),(B,M,K), N -> (B,C,N).  grouping.h:11-14 / grouping_kernel.cu:106-152."""
    grad_output = _cuda_f32(grad_output, "grad_output")
    index = _cuda_i64(index, "index")
    _require(grad_output.dim() == 4, "grad_output.dim() != 4")
    _require(index.dim() == 3, "index.dim() != 3")
    B, C, M, K = grad_output.shape
    _require(index.size(0) == B, "index.size(0) != batch_size")
    _require(index.size(1) == M, "index.size(1) != num_select")
    _require(index.size(2) == K, "index.size(2) != k")
    N = int(num_points)
    with torch.cuda.device(grad_output.device):
        grad_in = torch.empty((B, C, N), dtype=grad_output.dtype, device=grad_output.device)
        check(_fn("s4g_group_points_backward", grad_output)(ptr(grad_output), ptr(index), B, C, N, M, K, ptr(grad_in),
                                                stream_ptr(grad_output.device)), "group_points_backward")
    return grad_in


def point_search(query_xyz, key_xyz, num_neighbours):
    """(B,3,Nq),(B,3,Nk), 3 -> [index (B,Nq,3) int64, squared distance (B,Nq,3)].  interpolate.h:8-11."""
    query_xyz = _cuda_f32(query_xyz, "query_xyz")
    key_xyz = _cuda_f32(key_xyz, "key_xyz", query_xyz)
    B, _, Nq = query_xyz.shape
    _require(key_xyz.size(0) == B, "key_xyz.size(0) != batch_size")
    _require(query_xyz.size(1) == 3, "query_xyz.size(1) != 3")
    _require(key_xyz.size(1) == 3, "key_xyz.size(1) != 3")
    _require(int(num_neighbours) == 3, "num_neighbours != K")
    Nk = key_xyz.size(2)
    _require(Nk >= 3, "num_key < num_neighbours")
    with torch.cuda.device(query_xyz.device):
        index = torch.empty((B, Nq, 3), dtype=torch.int64, device=query_xyz.device)
        distance = torch.empty((B, Nq, 3), dtype=query_xyz.dtype, device=query_xyz.device)
        check(_fn("s4g_point_search", query_xyz)(ptr(query_xyz), ptr(key_xyz), B, Nq, Nk, 3, ptr(index), ptr(distance),
                                       stream_ptr(query_xyz.device)), "point_search")
    return [index, distance]


def interpolate_forward(input, index, weight):
    """(B,C,Nk),(B,Nq,3),(B,Nq,3) -> (B,C,Nq).  interpolate.h:13-16."""
    input = _cuda_f32(input, "input")
    index = _cuda_i64(index, "index")
    weight = _cuda_f32(weight, "weight", input)
    B, C, Nk = input.shape
    Nq = index.size(1)
    _require(index.size(0) == B, "index.size(0) != batch_size")
    _require(index.size(2) == 3, "index.size(2) != K")
    _require(weight.size(0) == B, "weight.size(0) != batch_size")
    _require(weight.size(1) == Nq, "weight.size(1) != num_select")
    _require(weight.size(2) == 3, "weight.size(2) != K")
    with torch.cuda.device(input.device):
        out = torch.empty((B, C, Nq), dtype=input.dtype, device=input.device)
        check(_fn("s4g_interpolate_forward", input)(ptr(input), ptr(index), ptr(weight), B, C, Nk, Nq, ptr(out),
                                              stream_ptr(input.device)), "interpolate_forward")
    return out


def interpolate_backward(grad_output, index, weight, num_inst):
    """(B,C,Nq),(B,Nq,3),(B,Nq,3), Nk -> (B,C,Nk).  interpolate.h:18-22."""
    grad_output = _cuda_f32(grad_output, "grad_output")
    index = _cuda_i64(index, "index")
    weight = _cuda_f32(weight, "weight", grad_output)
    B, C, Nq = grad_output.shape
    _require(index.size(0) == B, "index.size(0) != batch_size")
    _require(index.size(2) == 3, "index.size(2) != K")
    _require(weight.size(0) == B, "weight.size(0) != batch_size")
    _require(weight.size(1) == Nq, "weight.size(1) != num_select")
    _require(weight.size(2) == 3, "weight.size(2) != K")
    Nk = int(num_inst)
    with torch.cuda.device(grad_output.device):
        grad_in = torch.empty((B, C, Nk), dtype=grad_output.dtype, device=grad_output.device)
        check(_fn("s4g_interpolate_backward", grad_output)(ptr(grad_output), ptr(index), ptr(weight), B, C, Nk, Nq,
                                                           ptr(grad_in), stream_ptr(grad_output.device)),
              "interpolate_backward")
    return grad_in
