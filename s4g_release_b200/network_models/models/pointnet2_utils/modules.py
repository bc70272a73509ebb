"""PointNet++ building blocks with the reference's class names and constructor arguments
(inference/grasp_proposal/network_models/models/pointnet2_utils/modules.py): FarthestPointSampler
(:9), QueryGrouper (:30), FeatureInterpolator (:96), PointNetSAModule (:174), PointnetFPModule (:478), and the
variants the sibling models use: PointNetSAAvgModule (:253), PointNetSAModuleMSG (:330),
EdgeQueryGrouper (:63), EdgeFeatureInterpolator (:135), EdgeSAModule (:406), EdgeFPModule (:515).

These are the autograd-capable (training / drop-in) forms built on the sm_100a ops; eval-mode
inference of the whole PN2_CLS / PN2 network goes through the fused kernels (see models/PointNet2_tcls.py).
"""
import torch
from torch import nn

from ...functions.gather_knn import gather_knn
from ...nn_utils.mlp import SharedMLP
from . import functions as _F


class FarthestPointSampler(nn.Module):
    def __init__(self, num_centroids):
        super().__init__()
        self.num_centroids = num_centroids

    def forward(self, points):
        with torch.no_grad():
            return _F.farthest_point_sample(points, self.num_centroids)

    def extra_repr(self):
        return 'num_centroids={:d}'.format(self.num_centroids)


class QueryGrouper(nn.Module):
    def __init__(self, radius, num_neighbours):
        super().__init__()
        assert radius > 0.0 and num_neighbours > 0
        self.radius = radius
        self.num_neighbours = num_neighbours

    def forward(self, new_xyz, xyz, feature, use_xyz):
        with torch.no_grad():
            index, _ = _F.ball_query(xyz, new_xyz, self.radius, self.num_neighbours)
        # neighbour coordinates relative to their centroid: (B, 3, M, K)
        group_xyz = _F.group_points(xyz, index) - new_xyz.unsqueeze(-1)
        if feature is None:
            return group_xyz, group_xyz
        group_feature = _F.group_points(feature, index)
        if use_xyz:
            group_feature = torch.cat([group_xyz, group_feature], dim=1)  # xyz channels first
        return group_feature, group_xyz

    def extra_repr(self):
        return 'radius={}, num_neighbours={}'.format(self.radius, self.num_neighbours)


class FeatureInterpolator(nn.Module):
    def __init__(self, num_neighbors, eps=1e-10):
        super().__init__()
        self.num_neighbors = num_neighbors
        self._eps = eps

    def forward(self, dense_xyz, sparse_xyz, dense_feature, sparse_feature):
        with torch.no_grad():
            index, distance = _F.search_nn_distance(dense_xyz, sparse_xyz, self.num_neighbors)
            inv_distance = 1.0 / torch.clamp(distance, min=self._eps)  # on the squared distance
            weight = inv_distance / torch.sum(inv_distance, dim=2, keepdim=True)
        interpolated = _F.feature_interpolate(sparse_feature, index, weight)
        if dense_feature is None:
            return interpolated
        return torch.cat([interpolated, dense_feature], dim=1)  # interpolated channels first

    def extra_repr(self):
        return 'num_neighbours={:d}, eps={}'.format(self.num_neighbors, self._eps)


class PointNetSAModule(nn.Module):
    """Set abstraction: FPS -> ball query -> group -> SharedMLP(ndim=2) -> max over neighbours."""

    def __init__(self, in_channels, mlp_channels, num_centroids, radius, num_neighbours, use_xyz):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = mlp_channels[-1]
        self.num_centroids = num_centroids
        self.use_xyz = use_xyz
        self.mlp = SharedMLP(in_channels + (3 if use_xyz else 0), mlp_channels, ndim=2, bn=True)
        self.sampler = FarthestPointSampler(num_centroids) if num_centroids > 0 else None
        if num_neighbours < 0:
            assert radius < 0.0
            self.grouper = None
        else:
            assert num_neighbours > 0 and radius > 0.0
            self.grouper = QueryGrouper(radius, num_neighbours)

    def forward(self, xyz, feature=None):
        if self.num_centroids == 0:  # one global group centred at the origin
            assert self.grouper is None
            new_xyz = xyz.new_zeros(xyz.size(0), 3, 1)
            group_feature = feature.unsqueeze(2)
            if self.use_xyz:
                group_feature = torch.cat([xyz.unsqueeze(2), group_feature], dim=1)
        else:
            if self.num_centroids == -1:
                new_xyz = xyz
            else:
                new_xyz = _F.gather_points(xyz, self.sampler(xyz))
            group_feature, _ = self.grouper(new_xyz, xyz, feature, use_xyz=self.use_xyz)
        new_feature = self.mlp(group_feature)
        return new_xyz, torch.max(new_feature, 3)[0]

    def init_weights(self, init_fn=None):
        self.mlp.init_weights(init_fn)

    def extra_repr(self):
        return 'num_centroids={:d}, use_xyz={}'.format(self.num_centroids, self.use_xyz)


class PointnetFPModule(nn.Module):
    """Feature propagation: 3-NN inverse-distance interpolation -> concat -> SharedMLP(ndim=1)."""

    def __init__(self, in_channels, mlp_channels, num_neighbors):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = mlp_channels[-1]
        self.mlp = SharedMLP(in_channels, mlp_channels, ndim=1, bn=True)
        if num_neighbors == 0:
            self.interpolator = None
        elif num_neighbors == 3:
            self.interpolator = FeatureInterpolator(num_neighbors)
        else:
            raise ValueError('Expected value 1 or 3, but {} given.'.format(num_neighbors))

    def forward(self, dense_xyz, sparse_xyz, dense_feature, sparse_feature):
        if self.interpolator is None:
            assert sparse_xyz.size(2) == 1 and sparse_feature.size(2) == 1
            expanded = sparse_feature.expand(-1, -1, dense_xyz.size(2))
            new_feature = torch.cat([expanded, dense_feature], dim=1)
        else:
            new_feature = self.interpolator(dense_xyz, sparse_xyz, dense_feature, sparse_feature)
        return self.mlp(new_feature)

    def init_weights(self, init_fn=None):
        self.mlp.init_weights(init_fn)


class PointNetSAAvgModule(PointNetSAModule):
    """Set abstraction with MEAN pooling over the neighbours instead of max (reference modules.py:253-327); same
    parameters and state_dict keys as PointNetSAModule."""

    def forward(self, xyz, feature=None):
        if self.num_centroids == 0:
            assert self.grouper is None
            new_xyz = xyz.new_zeros(xyz.size(0), 3, 1)
            group_feature = feature.unsqueeze(2)
            if self.use_xyz:
                group_feature = torch.cat([xyz.unsqueeze(2), group_feature], dim=1)
        else:
            new_xyz = xyz if self.num_centroids == -1 else _F.gather_points(xyz, self.sampler(xyz))
            group_feature, _ = self.grouper(new_xyz, xyz, feature, use_xyz=self.use_xyz)
        return new_xyz, torch.mean(self.mlp(group_feature), 3)


class PointNetSAModuleMSG(nn.Module):
    """Multi-scale grouping (reference modules.py:330-401): ONE farthest-point sample, then per scale its own ball query
    (radius_list[i], num_neighbours_list[i]) -> group -> SharedMLP -> max; the scales' features are concatenated in
    order.  num_centroids == -1 keeps every point as a centroid."""

    def __init__(self, in_channels, mlp_channels_list, num_centroids, radius_list, num_neighbours_list, use_xyz):
        super().__init__()
        n_scales = len(mlp_channels_list)
        assert len(radius_list) == n_scales and len(num_neighbours_list) == n_scales
        self.in_channels = in_channels
        self.out_channels = sum(ch[-1] for ch in mlp_channels_list)
        self.num_centroids = num_centroids
        self.use_xyz = use_xyz
        self.mlp = nn.ModuleList()
        if num_centroids == -1:
            self.sampler = None
        else:
            assert num_centroids > 0
            self.sampler = FarthestPointSampler(num_centroids)
        self.grouper = nn.ModuleList()
        c = in_channels + (3 if use_xyz else 0)
        for ch, r, k in zip(mlp_channels_list, radius_list, num_neighbours_list):  # mlp / grouper interleaved like the reference
            self.mlp.append(SharedMLP(c, ch, ndim=2, bn=True))
            self.grouper.append(QueryGrouper(r, k))

    def forward(self, xyz, feature=None):
        new_xyz = _F.gather_points(xyz, self.sampler(xyz)) if self.num_centroids > 0 else xyz
        scales = []
        for mlp, grouper in zip(self.mlp, self.grouper):
            group_feature, _ = grouper(new_xyz, xyz, feature, use_xyz=self.use_xyz)
            scales.append(torch.max(mlp(group_feature), 3)[0])
        return new_xyz, torch.cat(scales, dim=1)

    def init_weights(self, init_fn=None):
        for mlp in self.mlp:
            mlp.init_weights(init_fn)

    def extra_repr(self):
        return 'num_centroids={:d}, use_xyz={}'.format(self.num_centroids, self.use_xyz)


# ---------------------------------------------------------------------------- EdgeConv variants
class EdgeQueryGrouper(nn.Module):
    """Ball-query grouping with edge features: [rel. xyz | neighbour feature | neighbour - centroid feature]."""

    def __init__(self, radius, num_neighbours):
        super().__init__()
        assert radius > 0.0 and num_neighbours > 0
        self.radius = radius
        self.num_neighbours = num_neighbours

    def forward(self, new_xyz, xyz, centroid_feature, feature, use_xyz):
        with torch.no_grad():
            index, _ = _F.ball_query(xyz, new_xyz, self.radius, self.num_neighbours)
        group_xyz = _F.group_points(xyz, index) - new_xyz.unsqueeze(-1)
        if feature is None:
            return group_xyz, group_xyz
        neighbour = _F.group_points(feature, index)
        parts = [neighbour, neighbour - centroid_feature.unsqueeze(-1)]
        if use_xyz:
            parts.insert(0, group_xyz)
        return torch.cat(parts, dim=1), group_xyz

    def extra_repr(self):
        return 'radius={}, num_neighbours={}'.format(self.radius, self.num_neighbours)


class EdgeFeatureInterpolator(nn.Module):
    """Per dense point and each of its K nearest sparse points: [interpolated | neighbour - interpolated | dense]
    -> (B, 2 C2 + C1, N1, K).  The neighbour gather carries no gradient, as in the reference (:158)."""

    def __init__(self, num_neighbors, eps=1e-10):
        super().__init__()
        self.num_neighbors = num_neighbors
        self._eps = eps

    def forward(self, dense_xyz, sparse_xyz, dense_feature, sparse_feature):
        with torch.no_grad():
            index, distance = _F.search_nn_distance(dense_xyz, sparse_xyz, self.num_neighbors)
            inv_distance = 1.0 / torch.clamp(distance, min=self._eps)
            weight = inv_distance / torch.sum(inv_distance, dim=2, keepdim=True)
            neighbour = gather_knn(sparse_feature, index)
        interpolated = _F.feature_interpolate(sparse_feature, index, weight)
        interpolated = interpolated.unsqueeze(-1).expand(-1, -1, -1, self.num_neighbors)
        parts = [interpolated, neighbour - interpolated]
        if dense_feature is not None:
            parts.append(dense_feature.unsqueeze(-1).expand(-1, -1, -1, self.num_neighbors))
        return torch.cat(parts, dim=1)


class EdgeSAModule(nn.Module):
    """Set abstraction on edge features (the MLP sees 2 x in_channels unless the level is the global one)."""

    def __init__(self, in_channels, mlp_channels, num_centroids, radius, num_neighbours, use_xyz):
        super().__init__()
        if num_centroids != 0:
            in_channels *= 2
        self.out_channels = mlp_channels[-1]
        self.num_centroids = num_centroids
        self.use_xyz = use_xyz
        self.mlp = SharedMLP(in_channels + (3 if use_xyz else 0), mlp_channels, ndim=2, bn=True)
        self.sampler = FarthestPointSampler(num_centroids) if num_centroids > 0 else None
        if num_neighbours < 0:
            assert radius < 0.0
            self.grouper = None
        else:
            assert num_neighbours > 0 and radius > 0.0
            self.grouper = EdgeQueryGrouper(radius, num_neighbours)

    def forward(self, xyz, feature=None):
        if self.num_centroids == 0:
            assert self.grouper is None
            new_xyz = xyz.new_zeros(xyz.size(0), 3, 1)
            group_feature = feature.unsqueeze(2)
            if self.use_xyz:
                group_feature = torch.cat([xyz.unsqueeze(2), group_feature], dim=1)
        else:
            if self.num_centroids == -1:
                new_xyz, centroid_feature = xyz, feature
            else:
                index = self.sampler(xyz)
                new_xyz = _F.gather_points(xyz, index)
                centroid_feature = _F.gather_points(feature, index) if feature is not None else None
            group_feature, _ = self.grouper(new_xyz, xyz, centroid_feature, feature, use_xyz=self.use_xyz)
        return new_xyz, torch.max(self.mlp(group_feature), 3)[0]


class EdgeFPModule(nn.Module):
    """Feature propagation on edge features: SharedMLP(ndim=2) over the K interpolation neighbours, then their mean."""

    def __init__(self, in_channels, mlp_channels, num_neighbors):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = mlp_channels[-1]
        if num_neighbors == 0:
            self.interpolator = None
            self.mlp = SharedMLP(in_channels, mlp_channels, ndim=1, bn=True)
        elif num_neighbors == 3:
            self.interpolator = EdgeFeatureInterpolator(num_neighbors)
            self.mlp = SharedMLP(in_channels, mlp_channels, ndim=2, bn=True)
        else:
            raise ValueError('Expected value 1 or 3, but {} given.'.format(num_neighbors))

    def forward(self, dense_xyz, sparse_xyz, dense_feature, sparse_feature):
        if self.interpolator is None:
            assert sparse_xyz.size(2) == 1 and sparse_feature.size(2) == 1
            expanded = sparse_feature.expand(-1, -1, dense_xyz.size(2))
            return self.mlp(torch.cat([expanded, dense_feature], dim=1))
        return torch.mean(self.mlp(self.interpolator(dense_xyz, sparse_xyz, dense_feature, sparse_feature)), dim=-1)
