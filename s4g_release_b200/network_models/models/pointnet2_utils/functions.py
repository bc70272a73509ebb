"""Autograd surface over the sm_100a ops — mirrors the callables of the reference's
``pointnet2_utils/functions.py`` (:10 gather_points, :50 farthest_point_sample, :80 ball_query,
:109 group_points, :135 search_nn_distance, :174 feature_interpolate): same names, argument order,
shapes and gradient behaviour (index-producing ops are non-differentiable; the two gathers scatter
their gradient back with the matching backward kernel, reference :101-105 and :165-170).
"""
import torch

from . import pn2_ext


class _NoGrad(torch.autograd.Function):
    """Base for the ops whose outputs are indices / distances used without gradient."""

    @staticmethod
    def backward(ctx, *grad_outputs):
        return (None,) * ctx.n_inputs


class _FarthestPointSample(_NoGrad):
    @staticmethod
    def forward(ctx, points, num_centroids):
        ctx.n_inputs = 2
        index = pn2_ext.farthest_point_sample(points, num_centroids)
        ctx.mark_non_differentiable(index)
        return index


class _BallQuery(_NoGrad):
    @staticmethod
    def forward(ctx, points, centroids, radius, num_neighbours):
        ctx.n_inputs = 4
        index, count = pn2_ext.ball_query(points, centroids, radius, num_neighbours)
        ctx.mark_non_differentiable(index, count)
        return index, count


class _SearchNNDistance(_NoGrad):
    @staticmethod
    def forward(ctx, query_xyz, key_xyz, num_neighbors):
        ctx.n_inputs = 3
        index, distance = pn2_ext.point_search(query_xyz, key_xyz, num_neighbors)
        ctx.mark_non_differentiable(index, distance)
        return index, distance


class _GroupPoints(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, index):
        ctx.save_for_backward(index)
        ctx.num_points = points.size(2)
        return pn2_ext.group_points_forward(points, index)

    @staticmethod
    def backward(ctx, grad_output):
        (index,) = ctx.saved_tensors
        return pn2_ext.group_points_backward(grad_output.contiguous(), index, ctx.num_points), None


class _FeatureInterpolate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feature, index, weight):
        ctx.save_for_backward(index, weight)
        ctx.num_inst = feature.size(2)
        return pn2_ext.interpolate_forward(feature, index, weight)

    @staticmethod
    def backward(ctx, grad_output):
        index, weight = ctx.saved_tensors
        return pn2_ext.interpolate_backward(grad_output.contiguous(), index, weight, ctx.num_inst), None, None


def gather_points(points, index):
    """points (B,C,N), index (B,M) -> (B,C,M) — reference functions.py:10-25.  Without autograd on fp32 CUDA tensors
    this is one launch of the library's gather kernel (s4g_gather_points_f32: the index row is read once for all C
    channels); whenever a gradient may be needed it is torch.gather, differentiable like the reference's."""
    if (points.is_cuda and points.dtype == torch.float32 and index.dtype == torch.int64 and
            not (torch.is_grad_enabled() and points.requires_grad)):
        return pn2_ext.gather_points(points, index)
    return points.gather(2, index.unsqueeze(1).expand(points.size(0), points.size(1), index.size(1)))


farthest_point_sample = _FarthestPointSample.apply
ball_query = _BallQuery.apply
group_points = _GroupPoints.apply
search_nn_distance = _SearchNNDistance.apply
feature_interpolate = _FeatureInterpolate.apply
