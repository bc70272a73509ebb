"""PN2_CLS — the S4G single-shot grasp network (the pretrained "curvature_model").

Module surface of inference/grasp_proposal/network_models/models/PointNet2_tcls.py: same class names,
constructor arguments, dict-in / dict-out ``forward`` (:99-148), parameter creation order (so a seeded
default init is identical) and state_dict keys (``sa_modules.{i}.mlp.{j}.conv.weight`` …,
200 entries / 6 632 213 parameters for the shipped configuration) — reference checkpoints load with
``strict=True``.

Two execution paths share the parameters:
  * module path (``self.training`` or autograd enabled): the nn.Module stack of pointnet2_utils.modules on
    the sm_100a ops — used for training and as the drop-in;
  * fused path (eval + no_grad on CUDA): ``engine.FusedPointNet2`` runs the whole forward with the
    hand-written kernels (geometry in fp32, MLP chains on tcgen05 tensor cores).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..nn_utils.mlp import SharedMLP
from .pointnet2_utils.modules import PointNetSAModule, PointnetFPModule

# configs/curvature_model.yaml:11-22 + yacs defaults: the shipped PN2_CLS hyper-parameters
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


class PointNet2(nn.Module):
    _SA_MODULE = PointNetSAModule
    _FP_MODULE = PointnetFPModule
    _R_OUT, _T_OUT = 9, 4  # frame_R: flattened 3x3; frame_t: 4 approach-offset classes (the PN2 sibling: 6 and 3)

    def __init__(self,
                 score_classes,
                 num_centroids=(10240, 1024, 128, 0),
                 radius=(0.2, 0.3, 0.4, -1.0),
                 num_neighbours=(64, 64, 64, -1),
                 sa_channels=((32, 32, 64), (64, 64, 128), (128, 128, 256), (256, 512, 1024)),
                 fp_channels=((256, 256), (256, 128), (128, 128), (64, 64, 64)),
                 num_fp_neighbours=(0, 3, 3, 3),
                 seg_channels=(128,),
                 num_removal_directions=5,
                 dropout_prob=0.5):
        super().__init__()
        n_sa, n_fp = len(num_centroids), len(fp_channels)
        assert len(radius) == n_sa and len(num_neighbours) == n_sa and len(sa_channels) == n_sa
        assert n_sa == n_fp and len(num_fp_neighbours) == n_fp
        self.config = dict(score_classes=score_classes, num_centroids=tuple(num_centroids), radius=tuple(radius),
                           num_neighbours=tuple(num_neighbours), sa_channels=tuple(map(tuple, sa_channels)),
                           fp_channels=tuple(map(tuple, fp_channels)), num_fp_neighbours=tuple(num_fp_neighbours),
                           seg_channels=tuple(seg_channels), num_removal_directions=num_removal_directions,
                           dropout_prob=dropout_prob)

        # encoder: set abstraction levels (xyz is always concatenated: use_xyz=True)
        self.sa_modules = nn.ModuleList()
        c = 0
        for i in range(n_sa):
            self.sa_modules.append(self._SA_MODULE(in_channels=c, mlp_channels=sa_channels[i],
                                                   num_centroids=num_centroids[i], radius=radius[i],
                                                   num_neighbours=num_neighbours[i], use_xyz=True))
            c = sa_channels[i][-1]
        skip = [0] + [ch[-1] for ch in sa_channels]

        # decoder: feature propagation levels, coarse to fine, with skip links
        self.fp_modules = nn.ModuleList()
        c = skip[-1]
        for i in range(n_fp):
            self.fp_modules.append(self._FP_MODULE(in_channels=c + skip[-2 - i], mlp_channels=fp_channels[i],
                                                   num_neighbors=num_fp_neighbours[i]))
            c = fp_channels[i][-1]

        # four per-point heads
        self.mlp_seg = SharedMLP(c, seg_channels, ndim=1, dropout_prob=dropout_prob)
        self.seg_logit = nn.Conv1d(seg_channels[-1], score_classes, 1, bias=True)
        self.mlp_R = SharedMLP(c, seg_channels, ndim=1)
        self.R_logit = nn.Conv1d(seg_channels[-1], self._R_OUT, 1, bias=True)
        self.mlp_t = SharedMLP(c, seg_channels, ndim=1)
        self.t_logit = nn.Conv1d(seg_channels[-1], self._T_OUT, 1, bias=True)
        self.mlp_movable = SharedMLP(c, seg_channels, ndim=1, dropout_prob=dropout_prob)
        self.movable_logit = nn.Sequential(nn.Conv1d(seg_channels[-1], num_removal_directions, 1, bias=True),
                                           nn.Sigmoid())
        self._engine = None

    # ------------------------------------------------------------------ module (autograd) path
    def forward_modules(self, points):
        xyz, feature = points, None
        level_xyz, level_feature = [xyz], [feature]
        for sa in self.sa_modules:
            xyz, feature = sa(xyz, feature)
            level_xyz.append(xyz)
            level_feature.append(feature)
        sparse_xyz, sparse_feature = xyz, feature
        for i, fp in enumerate(self.fp_modules):
            dense_xyz, dense_feature = level_xyz[-2 - i], level_feature[-2 - i]
            sparse_feature = fp(dense_xyz, sparse_xyz, dense_feature, sparse_feature)
            sparse_xyz = dense_xyz
        return {
            "score": self.seg_logit(self.mlp_seg(sparse_feature)),
            "frame_R": self.R_logit(self.mlp_R(sparse_feature)),
            "frame_t": self.t_logit(self.mlp_t(sparse_feature)),
            "movable_logits": self.movable_logit(self.mlp_movable(sparse_feature)),
        }

    # ------------------------------------------------------------------ fused (inference) path
    def fused_engine(self, refresh=False):
        """The fused sm_100a inference engine bound to this module's current parameters."""
        if self._engine is None or refresh:
            from ...engine import FusedPointNet2
            self._engine = FusedPointNet2(self)
        return self._engine

    def train(self, mode=True):
        self._engine = None  # parameters may change: re-fold BN on the next eval forward
        return super().train(mode)

    def forward(self, data_batch, fused=None, host_out=None):
        """``host_out`` (fused path only): dict of pinned host tensors that receive the predictions while the later
        heads are still computing (engine.FusedPointNet2.forward)."""
        points = data_batch["scene_points"]
        if fused is None:
            fused = (not self.training) and (not torch.is_grad_enabled()) and points.is_cuda
        if fused:
            return self.fused_engine().forward(points, host_out=host_out)
        return self.forward_modules(points)

    def init_weights(self):
        pass


class PointNet2Loss(nn.Module):
    """Reference PointNet2_tcls.py:156-219: weighted CE on the score classes (class 0 down-weighted),
    L1 on the movable directions, flip-symmetric rotation MSE weighted by the ground-truth score (x5)
    and CE on the 4 approach-offset classes (x0.2).  The reference reads ``preds["scene_score_logits"]``
    although this model's forward emits ``"score"`` (:142,:163); both keys are accepted."""

    def __init__(self, label_smoothing=0, neg_weight=0.1):
        super().__init__()
        if label_smoothing > 0:
            raise NotImplementedError("label smoothing is 0.0 in the shipped PN2_CLS configuration")
        self.label_smoothing = label_smoothing
        self.neg_weight = neg_weight

    def forward(self, preds, labels):
        logits = preds["scene_score_logits"] if "scene_score_logits" in preds else preds["score"]
        weight = torch.ones(logits.shape[1], device=logits.device)
        weight[0] = self.neg_weight
        cls_loss = F.cross_entropy(logits, labels["scene_score_labels"], weight)
        mov_loss = F.l1_loss(preds["movable_logits"], labels["scene_movable_labels"])

        gt_R = labels["best_frame_R"]
        n = gt_R.shape[2]
        pred_R = preds["frame_R"][:, :, :n]
        flip = gt_R.new_tensor([1, -1, -1, 1, -1, -1, 1, -1, -1]).view(1, 9, 1)  # columns 1,2 negated
        loss_a = ((pred_R - gt_R) ** 2).mean(1)
        loss_b = ((pred_R - gt_R * flip) ** 2).mean(1)
        R_loss = (torch.minimum(loss_a, loss_b) * labels["scene_score"][:, :n]).mean() * 5.0
        t_loss = F.cross_entropy(preds["frame_t"][:, :, :n], labels["best_frame_t"]) * 0.2
        return {"cls_loss": cls_loss, "R_loss": R_loss, "t_loss": t_loss, "mov_loss": mov_loss}


def build_pointnet2_cls(cfg=None):
    """Reference build_pointnet2_cls (:270-290).  ``cfg`` may be a yacs-style node (MODEL.PN2.*, DATA.*)
    or None for the shipped curvature_model.yaml values."""
    if cfg is None:
        net = PointNet2(**PN2_CLS_CONFIG)
        return net, PointNet2Loss(label_smoothing=0.0, neg_weight=0.5), None
    pn2 = cfg.MODEL.PN2
    net = PointNet2(score_classes=cfg.DATA.SCORE_CLASSES, num_centroids=pn2.NUM_CENTROIDS, radius=pn2.RADIUS,
                    num_neighbours=pn2.NUM_NEIGHBOURS, sa_channels=pn2.SA_CHANNELS, fp_channels=pn2.FP_CHANNELS,
                    num_fp_neighbours=pn2.NUM_FP_NEIGHBOURS, seg_channels=pn2.SEG_CHANNELS,
                    num_removal_directions=cfg.DATA.NUM_REMOVAL_DIRECTIONS, dropout_prob=pn2.DROPOUT_PROB)
    return net, PointNet2Loss(label_smoothing=pn2.LABEL_SMOOTHING, neg_weight=pn2.NEG_WEIGHT), None
