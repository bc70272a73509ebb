"""ORACLE — TEST INFRASTRUCTURE ONLY (not the product path).

ctypes front-end of ``libpn2_oracle.so`` (oracle/pn2_oracle.c) that exposes the
seven functions of the reference's ``pn2_ext`` module
(inference/grasp_proposal/network_models/models/pointnet2_utils/csrc/main.cpp:7-13)
on **CPU** torch tensors, with the reference's signatures, channel-first shapes,
int64 indices and precondition errors.  It can be injected as
``sys.modules['…pointnet2_utils.pn2_ext']`` so the reference's own python
modules run on the host (tests/golden/make_golden.py does that), and it is the
checker the GPU parity tests compare against.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import it.
"""
import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libpn2_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)
_i64p = ctypes.POINTER(ctypes.c_int64)
_i64 = ctypes.c_int64


def build(force=False):
    """Compile the C restatement (gcc, a few seconds)."""
    src = os.path.join(_HERE, "pn2_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libpn2_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.pn2o_num_threads.restype = ctypes.c_int
    return _lib


def num_threads():
    return int(lib().pn2o_num_threads())


def _fp(t):
    return ctypes.cast(t.data_ptr(), ctypes.c_void_p)


def _fn(name, t):
    """the fp32 or fp64 instantiation (the reference dispatches both: AT_DISPATCH_FLOATING_TYPES)"""
    return getattr(lib(), name + ("_f64" if t.dtype == torch.float64 else ""))


def _ip(t):
    return ctypes.cast(t.data_ptr(), _i64p)


def _f32(t, name):
    if t.is_cuda:
        raise RuntimeError(f"{name}: the oracle runs on CPU tensors")
    if t.dtype not in (torch.float32, torch.float64):
        raise RuntimeError(f"{name} must be float32 or float64 in the oracle")
    return t.contiguous()


def _check(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def farthest_point_sample(points, num_centroids, keyed=False):
    """sampling.h:7-9 / sampling_kernel.cu:128-172.  (B,3,N) -> (B,M) int64."""
    points = _f32(points, "points")
    _check(points.dim() == 3 and points.size(1) == 3, "points.size(1) != 3")
    B, _, N = points.shape
    M = int(num_centroids)
    _check(M > 0, "num_centroids <= 0")
    _check(N >= M, "num_points < num_centroids")
    index = torch.zeros(B, M, dtype=torch.int64)
    fn = _fn("pn2o_farthest_point_sample_keyed" if keyed else "pn2o_farthest_point_sample", points)
    fn(_fp(points), _i64(B), _i64(N), _i64(M), _ip(index))
    return index


def ball_query(points, centroids, radius, num_neighbours):
    """ball_query.h:7-11 / ball_query_kernel.cu:89-133 -> [index (B,M,K), count (B,M)]."""
    points = _f32(points, "points")
    centroids = _f32(centroids, "centroids")
    _check(points.size(1) == 3, "points.size(1) != 3")
    _check(centroids.size(1) == 3, "centroids.size(1) != 3")
    B, _, N = points.shape
    M = centroids.size(2)
    K = int(num_neighbours)
    index = torch.zeros(B, M, K, dtype=torch.int64)
    count = torch.zeros(B, M, dtype=torch.int64)
    # the reference passes its `const float radius` converted to scalar_t (ball_query_kernel.cu:116-127)
    r = ctypes.c_double(ctypes.c_float(radius).value) if points.dtype == torch.float64 else ctypes.c_float(radius)
    _fn("pn2o_ball_query", points)(_fp(points), _fp(centroids), _i64(B), _i64(N), _i64(M), r, _i64(K), _ip(index),
                                   _ip(count))
    return [index, count]


def group_points_forward(input, index):
    """grouping.h:7-9 / grouping_kernel.cu:32-54.  (B,C,N),(B,M,K) -> (B,C,M,K)."""
    input = _f32(input, "input")
    index = index.contiguous()
    _check(input.dim() == 3, "input.dim() != 3")
    _check(index.dim() == 3, "index.dim() != 3")
    _check(index.size(0) == input.size(0), "index.size(0) != batch_size")
    B, C, N = input.shape
    _, M, K = index.shape
    out = torch.empty(B, C, M, K, dtype=input.dtype)
    _fn("pn2o_group_points_forward", input)(_fp(input), _ip(index), _i64(B), _i64(C), _i64(N), _i64(M), _i64(K), _fp(out))
    return out


def group_points_backward(grad_output, index, num_points):
    """grouping.h:11-14 / grouping_kernel.cu:106-152.  (B,C,M,K) -> (B,C,N)."""
    grad_output = _f32(grad_output, "grad_output")
    index = index.contiguous()
    _check(grad_output.dim() == 4, "grad_output.dim() != 4")
    B, C, M, K = grad_output.shape
    _check(tuple(index.shape) == (B, M, K), "index shape mismatch")
    grad_in = torch.empty(B, C, int(num_points), dtype=grad_output.dtype)
    _fn("pn2o_group_points_backward", grad_output)(_fp(grad_output), _ip(index), _i64(B), _i64(C), _i64(num_points),
                                     _i64(M), _i64(K), _fp(grad_in))
    return grad_in


def point_search(query_xyz, key_xyz, num_neighbours):
    """interpolate.h:8-11 / interpolate_kernel.cu:92-132 -> [index (B,Nq,3), dist² (B,Nq,3)]."""
    query_xyz = _f32(query_xyz, "query_xyz")
    key_xyz = _f32(key_xyz, "key_xyz")
    B, _, Nq = query_xyz.shape
    Nk = key_xyz.size(2)
    _check(key_xyz.size(0) == B, "key_xyz.size(0) != batch_size")
    _check(query_xyz.size(1) == 3 and key_xyz.size(1) == 3, "xyz.size(1) != 3")
    _check(int(num_neighbours) == 3, "num_neighbours != K")
    _check(Nk >= 3, "num_key < num_neighbours")
    index = torch.zeros(B, Nq, 3, dtype=torch.int64)
    dist = torch.zeros(B, Nq, 3, dtype=query_xyz.dtype)
    _fn("pn2o_point_search", query_xyz)(_fp(query_xyz), _fp(key_xyz), _i64(B), _i64(Nq), _i64(Nk), _ip(index), _fp(dist))
    return [index, dist]


def interpolate_forward(input, index, weight):
    """interpolate.h:13-16 / interpolate_kernel.cu:191-236.  (B,C,Nk) -> (B,C,Nq)."""
    input = _f32(input, "input")
    index = index.contiguous()
    weight = _f32(weight, "weight")
    B, C, Nk = input.shape
    Nq = index.size(1)
    _check(index.size(0) == B and index.size(2) == 3, "index shape mismatch")
    _check(tuple(weight.shape) == (B, Nq, 3), "weight shape mismatch")
    out = torch.empty(B, C, Nq, dtype=input.dtype)
    _fn("pn2o_interpolate_forward", input)(_fp(input), _ip(index), _fp(weight), _i64(B), _i64(C), _i64(Nk), _i64(Nq), _fp(out))
    return out


def interpolate_backward(grad_output, index, weight, num_inst):
    """interpolate.h:18-22 / interpolate_kernel.cu:296-341.  (B,C,Nq) -> (B,C,Nk)."""
    grad_output = _f32(grad_output, "grad_output")
    index = index.contiguous()
    weight = _f32(weight, "weight")
    B, C, Nq = grad_output.shape
    _check(index.size(0) == B and index.size(2) == 3, "index shape mismatch")
    _check(tuple(weight.shape) == (B, Nq, 3), "weight shape mismatch")
    grad_in = torch.empty(B, C, int(num_inst), dtype=grad_output.dtype)
    _fn("pn2o_interpolate_backward", grad_output)(_fp(grad_output), _ip(index), _fp(weight), _i64(B), _i64(C), _i64(num_inst),
                                    _i64(Nq), _fp(grad_in))
    return grad_in


def gather_points(points, index):
    """functions.py:10-25."""
    points = _f32(points, "points")
    index = index.contiguous()
    B, C, N = points.shape
    M = index.size(1)
    out = torch.empty(B, C, M, dtype=points.dtype)
    _fn("pn2o_gather_points", points)(_fp(points), _ip(index), _i64(B), _i64(C), _i64(N), _i64(M), _fp(out))
    return out
