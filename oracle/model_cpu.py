"""ORACLE — TEST INFRASTRUCTURE ONLY (not the product path).

Functional fp32 CPU restatement of the PN2_CLS forward of yzqin/s4g-release
(eval mode), driven by a reference-layout ``state_dict``.  It calls the C
restatement of the ``pn2_ext`` ops (oracle/pn2_ext_cpu.py) and plain torch-CPU
fp32 arithmetic for conv/BN/ReLU/max (the reference itself delegates those to
PyTorch: nn_utils/conv.py:24-36,64-76).

Pinned by tests/golden/make_golden.py: run in the build container it imports the
reference's own python modules (with pn2_ext_cpu injected) and checks this file
reproduces their output bit-for-bit before writing the fixtures.

Paths cited are relative to inference/grasp_proposal/.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import pn2_ext_cpu as ops

# configs/curvature_model.yaml:11-22 (+ yacs defaults)
PN2_CLS_CONFIG = dict(
    score_classes=3,
    num_centroids=(5120, 1024, 256),
    radius=(0.02, 0.08, 0.32),
    num_neighbours=(64, 64, 64),
    sa_channels=((128, 128, 256), (256, 256, 512), (512, 512, 1024)),
    fp_channels=((1024, 1024), (512, 512), (256, 256, 256)),
    num_fp_neighbours=(3, 3, 3),
    seg_channels=(512, 256, 256, 128),
    num_removal_directions=5,
    dropout_prob=0.5,
)
NUM_INPUT = 25600
BN_EPS = 1e-5  # torch default, nn_utils/conv.py:25,65


def _conv_bn_relu(x, sd, prefix):
    """nn_utils/conv.py:30-36 (Conv1d) / :70-76 (Conv2d) in eval mode: bias-free 1x1 conv ->
    BatchNorm (running statistics, eps 1e-5) -> ReLU, with the same torch functional ops the
    reference's nn.Conv*/nn.BatchNorm* modules dispatch to."""
    w = sd[prefix + ".conv.weight"]
    conv = F.conv2d if x.dim() == 4 else F.conv1d
    y = conv(x, w)
    y = F.batch_norm(y, sd[prefix + ".bn.running_mean"], sd[prefix + ".bn.running_var"],
                     sd[prefix + ".bn.weight"], sd[prefix + ".bn.bias"], False, 0.1, BN_EPS)
    return F.relu(y)


def shared_mlp(x, sd, prefix, n_layers):
    """nn_utils/mlp.py:95-106 (eval: no dropout)."""
    for j in range(n_layers):
        x = _conv_bn_relu(x, sd, f"{prefix}.{j}")
    return x


def sa_module(xyz, feature, sd, prefix, n_layers, num_centroids, radius, num_neighbours, trace=None):
    """models/pointnet2_utils/modules.py:208-244 + QueryGrouper :37-54."""
    index = ops.farthest_point_sample(xyz, num_centroids)
    new_xyz = ops.gather_points(xyz, index)
    nbr, count = ops.ball_query(xyz, new_xyz, radius, num_neighbours)
    group_xyz = ops.group_points_forward(xyz, nbr)
    group_xyz = group_xyz - new_xyz.unsqueeze(-1)
    if feature is not None:
        group_feature = ops.group_points_forward(feature, nbr)
        group_feature = torch.cat([group_xyz, group_feature], dim=1)
    else:
        group_feature = group_xyz
    new_feature = shared_mlp(group_feature, sd, prefix + ".mlp", n_layers)
    new_feature = new_feature.max(dim=3)[0]
    if trace is not None:
        trace.append(dict(fps_index=index, ball_index=nbr, ball_count=count, new_xyz=new_xyz,
                          new_feature=new_feature))
    return new_xyz, new_feature


def interpolation_weights(distance, eps=1e-10):
    """modules.py:115-120 — on the SQUARED distance the op returns."""
    inv = 1.0 / torch.clamp(distance, min=eps)
    norm = torch.sum(inv, dim=2, keepdim=True)
    return inv / norm


def fp_module(dense_xyz, sparse_xyz, dense_feature, sparse_feature, sd, prefix, n_layers, trace=None):
    """modules.py:498-507 + FeatureInterpolator :102-129 (num_neighbors = 3)."""
    index, distance = ops.point_search(dense_xyz, sparse_xyz, 3)
    weight = interpolation_weights(distance)
    interpolated = ops.interpolate_forward(sparse_feature, index, weight)
    if dense_feature is not None:
        new_feature = torch.cat([interpolated, dense_feature], dim=1)
    else:
        new_feature = interpolated
    out = shared_mlp(new_feature, sd, prefix + ".mlp", n_layers)
    if trace is not None:
        trace.append(dict(nn_index=index, nn_dist=distance, weight=weight, fp_feature=out))
    return out


def _logit(x, sd, prefix):
    """nn.Conv1d(C, n, 1, bias=True): models/PointNet2_tcls.py:84,87,90,93."""
    return F.conv1d(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def pointnet2_forward(points, sd, cfg=PN2_CLS_CONFIG, trace=None):
    """models/PointNet2_tcls.py:99-148.  points (B,3,N) fp32 CPU -> dict of 4 tensors."""
    sd = {k: v.detach().to(torch.float32).cpu() for k, v in sd.items() if v.is_floating_point()}
    xyz, feature = points, None
    inter_xyz, inter_feature = [xyz], [feature]
    sa_trace = [] if trace is not None else None
    fp_trace = [] if trace is not None else None
    for i in range(len(cfg["num_centroids"])):
        xyz, feature = sa_module(xyz, feature, sd, f"sa_modules.{i}", len(cfg["sa_channels"][i]),
                                 cfg["num_centroids"][i], cfg["radius"][i], cfg["num_neighbours"][i], sa_trace)
        inter_xyz.append(xyz)
        inter_feature.append(feature)
    sparse_xyz, sparse_feature = xyz, feature
    for i in range(len(cfg["fp_channels"])):
        dense_xyz, dense_feature = inter_xyz[-2 - i], inter_feature[-2 - i]
        sparse_feature = fp_module(dense_xyz, sparse_xyz, dense_feature, sparse_feature, sd, f"fp_modules.{i}",
                                   len(cfg["fp_channels"][i]), fp_trace)
        sparse_xyz = dense_xyz
    n_seg = len(cfg["seg_channels"])
    x = shared_mlp(sparse_feature, sd, "mlp_seg", n_seg)
    logits = _logit(x, sd, "seg_logit")
    R = _logit(shared_mlp(sparse_feature, sd, "mlp_R", n_seg), sd, "R_logit")
    t = _logit(shared_mlp(sparse_feature, sd, "mlp_t", n_seg), sd, "t_logit")
    mov = torch.sigmoid(_logit(shared_mlp(sparse_feature, sd, "mlp_movable", n_seg), sd, "movable_logit.0"))
    if trace is not None:
        trace["sa"] = sa_trace
        trace["fp"] = fp_trace
        trace["point_feature"] = sparse_feature
    return {"score": logits, "frame_R": R, "frame_t": t, "movable_logits": mov}


# --------------------------------------------------------------------------- #
# Post-processing: grasp_detector.py:124-185 (numpy, float64 after .cpu())
# --------------------------------------------------------------------------- #
# grasp_detector.py:26-27
REAL2TRAIN = np.array([[0, 1, 0, 0], [1, 0, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1]], dtype=np.float64)
TRAIN2REAL = np.linalg.inv(REAL2TRAIN)
# configs/real_world_config.py:21-24
CAMERA2BASE = np.array([[-0.00377177, 0.54720216, -0.83699198, 0.766],
                        [0.99981506, -0.01372054, -0.01347562, -0.276],
                        [-0.01885787, -0.83688801, -0.54704921, 0.62],
                        [0., 0., 0., 1.]])
T_SCORE = np.array([0.08, 0.06, 0.04, 0.02])  # grasp_detector.py:177


def orthogonalization(batch_rotation, batch_translation):
    """grasp_detector.py:124-135 (Gram-Schmidt on columns 0,1; z = x × y)."""
    x = batch_rotation[:, :, 0]
    x = x / np.linalg.norm(x, axis=1, keepdims=True)
    y = batch_rotation[:, :, 1]
    y = y - np.sum(x * y, axis=1, keepdims=True) * x
    y = y / np.linalg.norm(y, axis=1, keepdims=True)
    z = np.cross(x, y)
    mat44 = np.tile(np.eye(4), [batch_rotation.shape[0], 1, 1])
    mat44[:, :3, :3] = np.stack([x, y, z], axis=2)
    mat44[:, :3, 3] = batch_translation
    return mat44


def grasp_scores(score_logits):
    """grasp_detector.py:142-145 for one scene: (C, N) logits -> (N,) fp64 expected score."""
    all_scores = torch.softmax(score_logits, dim=0).detach().cpu().numpy()
    score_classes = all_scores.shape[0]
    score_value = np.linspace(0, 1, score_classes + 1)[1:][:, np.newaxis]
    return np.sum(score_value * all_scores, axis=0)


def post_processing(points_array, predictions, score_threshold=0.7, vertical_degree_threshold=0.2, all_scores=None,
                    return_index=False):
    """grasp_detector.py:137-185 for batch element 0, including the behaviour that
    ``frame_R`` is indexed with positions *within* the filtered set (:153).  ``all_scores`` (optional)
    replaces the first step so a test can check the index logic exactly on another implementation's scores."""
    if all_scores is None:
        all_scores = grasp_scores(predictions["score"][0])
    high_score_index = np.nonzero(all_scores > score_threshold)[0]
    index_high2low = np.argsort(all_scores[high_score_index])[::-1]
    rotation = predictions["frame_R"][0].detach().cpu().numpy()[:, index_high2low]
    rotation = rotation.transpose(0, 1).reshape([-1, 3, 3])
    x_direction = -CAMERA2BASE[:3, :3] @ TRAIN2REAL[:3, :3] @ rotation[:, :, 0].T
    vertical_direction = np.array([[0, 0, 1]], dtype=np.float32)
    vertical_degree = np.sum(x_direction.T * vertical_direction, axis=1, keepdims=False)
    index_good_direction = np.nonzero(vertical_degree > vertical_degree_threshold)[0]
    valid_index = high_score_index[index_good_direction]
    if points_array.shape[0] == 3:
        points_array = points_array.T
    points = points_array[valid_index, :]
    rotation = rotation[index_good_direction, :, :]
    translation = torch.softmax(predictions["frame_t"][0][:, valid_index], dim=0)
    translation = translation.transpose(0, 1).detach().cpu().numpy()
    scores = all_scores[valid_index]
    global_translation = -(translation * T_SCORE[np.newaxis, :]).sum(1, keepdims=True) * rotation[:, :, 0] + points
    global_mat44 = orthogonalization(rotation, global_translation)
    global_mat44 = np.matmul(TRAIN2REAL[np.newaxis, :, :], global_mat44)
    if return_index:
        return global_mat44, scores, valid_index, index_good_direction
    return global_mat44, scores


# --------------------------------------------------------------------------- #
# Collision check, importance sampling, de-duplication (CPU restatements; test infrastructure only)
# --------------------------------------------------------------------------- #
# configs/gripper_config.py:9-21, configs/processing_config.py:19,37-40
GRIPPER = dict(half_bottom_width=0.057, bottom_length=0.16, finger_width=0.023, half_hand_thickness=0.012,
               finger_length=0.09, back_collision_margin=0.0, back_collision_threshold=10 * np.sqrt(8),
               finger_collision_threshold=10)
GRIPPER["half_bottom_space"] = GRIPPER["half_bottom_width"] - GRIPPER["finger_width"]


def batch_transformation_inv(poses):
    """utils/math_utils.py:27-40 on fp32 copies of the (n,4,4) poses."""
    t = torch.as_tensor(poses, dtype=torch.float32)
    out = torch.eye(4, dtype=torch.float32).unsqueeze(0).repeat(t.shape[0], 1, 1)
    out[:, :3, :3] = t[:, :3, :3].transpose(1, 2)
    out[:, :3, 3:] = torch.bmm(-t[:, :3, :3].transpose(1, 2), t[:, :3, 3:])
    return out


def view_non_collision(global2local, cloud_homo, g=GRIPPER, return_counts=False):
    """cloud_processor/view_collision_checker.py:37-65.  cloud_homo: (4, n) fp32."""
    local = torch.matmul(global2local, cloud_homo)
    close = (local[0] < g["finger_length"]) & (local[0] > -g["bottom_length"])
    pts = local[:, close][0:3]
    zc = (pts[2] < g["half_hand_thickness"]) & (pts[2] > -g["half_hand_thickness"])
    back = (pts[1] < g["half_bottom_width"]) & (pts[1] > -g["half_bottom_width"]) & \
           (pts[0] < -g["back_collision_margin"]) & zc
    left = (pts[1] < g["half_bottom_width"]) & (pts[1] > g["half_bottom_space"])
    right = (pts[1] > -g["half_bottom_width"]) & (pts[1] < -g["half_bottom_space"])
    finger = zc & (left | right)
    n_back, n_finger = int(back.sum()), int(finger.sum())
    ok = not (n_back > g["back_collision_threshold"]) and not (n_finger > g["finger_collision_threshold"])
    return (ok, n_back, n_finger) if return_counts else ok


def collision_filter(poses, cloud_n3):
    """grasp_detector.py:214-224: indices of the collision-free poses, plus the per-pose point counts."""
    cloud = torch.as_tensor(cloud_n3, dtype=torch.float32)
    homo = torch.cat([cloud.t(), torch.ones(1, cloud.shape[0])], dim=0)
    inv = batch_transformation_inv(poses)
    res = [view_non_collision(inv[i], homo, return_counts=True) for i in range(inv.shape[0])]
    ok = np.array([r[0] for r in res], dtype=bool)
    counts = np.array([[r[1], r[2]] for r in res], dtype=np.int64).reshape(-1, 2)
    return np.nonzero(ok)[0], counts


def importance_sampling(scores, sorted_uniform):
    """grasp_detector.py:235-246 with the random numbers passed in (np.sort(np.random.rand(k)))."""
    cum = np.cumsum(np.exp(5 * np.asarray(scores, dtype=np.float64)))
    out, index = [], 0
    for u in sorted_uniform:
        target = u * cum[-1]
        while cum[index] < target:
            index += 1
        out.append(index)
    return np.array(out, dtype=np.int64)


def translation_nms(poses, scores, min_dist):
    """Greedy de-duplication in descending score order (ties: lower index first): a pose is dropped when
    the L1 distance of its translation to a kept pose is < min_dist — the check sketched, commented out,
    at utils/file_logger_cls.py:220-225.  The reference ships no NMS (README.md:58); this is OUR definition."""
    order = np.argsort(-np.asarray(scores, dtype=np.float64), kind="stable")
    kept = []
    for c in order:
        t = poses[c, :3, 3]
        if all(np.abs(poses[k, :3, 3] - t).sum() >= min_dist for k in kept):
            kept.append(int(c))
    return np.array(kept, dtype=np.int64)


# --------------------------------------------------------------------------- #
# Pre-processing (CPU restatement; test infrastructure only)
# --------------------------------------------------------------------------- #
def transform_numpy_points(cloud_array, transformation_matrix):
    """utils/math_utils.py:20-24."""
    homo = np.concatenate([cloud_array, np.ones([1, cloud_array.shape[1]])], axis=0)
    return (transformation_matrix @ homo)[:3, :]


def pre_processing(cloud_array, random_index):
    """grasp_detector.py:92-105 as it BEHAVES: voxelize() / remove_outliers() discard open3d's return values
    (cloud_processor.py:31-42), so the cloud goes straight to _REAL2TRAIN and the random sub-sample (:86-91);
    the result is what eval() turns into the float32 network input (:113)."""
    points = transform_numpy_points(np.asarray(cloud_array), REAL2TRAIN.astype(np.int64))
    return points[:, np.asarray(random_index)].astype(np.float32)


def voxel_down_sample(cloud_3n, voxel_size):
    """OUR definition of the intended voxel filter (csrc/preprocess.cu): per-voxel mean, ascending (z,y,x) cell order,
    grid anchored at min - voxel/2 like open3d."""
    p = np.asarray(cloud_3n, dtype=np.float32)
    origin = (p.min(axis=1) - np.float32(0.5 * voxel_size)).astype(np.float32)
    extent = p.max(axis=1) - origin
    dims = np.maximum(np.floor(extent / voxel_size).astype(np.int64) + 1, 1)
    inv = np.float32(1.0) / np.float32(voxel_size)
    cell = np.floor((p - origin[:, None]) * inv).astype(np.int64)
    cell = np.minimum(np.maximum(cell, 0), (dims - 1)[:, None])
    key = (cell[2] * dims[1] + cell[1]) * dims[0] + cell[0]
    order = np.argsort(key, kind="stable")
    uniq, start = np.unique(key[order], return_index=True)
    out = np.zeros((3, len(uniq)), dtype=np.float32)
    bounds = list(start) + [len(key)]
    for v in range(len(uniq)):
        sel = order[bounds[v]:bounds[v + 1]]
        out[:, v] = (p[:, sel].astype(np.float64).sum(axis=1) / len(sel)).astype(np.float32)
    return out


def radius_outlier_mask(cloud_3n, nb_points, radius):
    """Keep points with more than nb_points points (itself included) within radius — fp32 squared distances in the
    ops' rounding order, strict '<' against radius^2 (the ball-query rule)."""
    from oracle import pn2_ext_cpu
    p = torch.as_tensor(np.asarray(cloud_3n, dtype=np.float32)).unsqueeze(0)
    _, count = pn2_ext_cpu.ball_query(p, p, radius, nb_points + 1)
    return (count[0] > nb_points).numpy()
