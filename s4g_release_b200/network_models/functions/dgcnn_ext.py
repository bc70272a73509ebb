"""Drop-in for the reference's compiled ``dgcnn_ext`` module (network_models/functions/csrc/gather_knn_kernel.cu:27-153,
bound in functions/gather_knn.py:4-7): the k-NN feature gather of the EdgeConv modules and its backward.

gather_knn_forward(feature (B,C,N), index (B,N,K) int64) -> (B,C,N,K) is ``group_points`` with one group per point, and
the backward is the same scatter-add, so both map onto the sm_100a grouping kernels of libs4g_b200.so
(s4g_group_points_forward/backward_f32|f64).  float32 and float64 like the reference's AT_DISPATCH."""
from ..models.pointnet2_utils import pn2_ext


def gather_knn_forward(feature, index):
    # (the reference leaves index.size(1) == N unchecked, gather_knn_kernel.cu:41, and EdgeFeatureInterpolator relies
    # on it: the queries are the DENSE points, the gathered features the sparse ones)
    if index.dim() != 3 or feature.dim() != 3 or index.size(0) != feature.size(0):
        raise RuntimeError("gather_knn_forward: feature (B,C,N) and index (B,M,K) expected")
    return pn2_ext.group_points_forward(feature, index)


def gather_knn_backward(grad_output, index):
    return pn2_ext.group_points_backward(grad_output, index, index.size(1))
