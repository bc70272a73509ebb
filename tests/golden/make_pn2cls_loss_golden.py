"""Golden fixture for the PN2_CLS loss and metric from the REFERENCE's own classes (build container only):
network_models/models/PointNet2_tcls.py::PointNet2Loss (label smoothing off and on) and ::PointNet2Metric, called on
seeded predictions / labels of BASELINE config 4's label layout.  -> tests/golden/pn2cls_loss.npz"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_golden import import_reference  # noqa: E402

import_reference()
from grasp_proposal.network_models.models.PointNet2_tcls import PointNet2Loss, PointNet2Metric  # noqa: E402

rs = np.random.RandomState(11)
B, N, n = 3, 512, 300
preds = {"scene_score_logits": torch.from_numpy(rs.randn(B, 3, N).astype(np.float32) * 2),
         "frame_R": torch.from_numpy((np.linalg.qr(rs.randn(B, N, 3, 3))[0] + 0.2 * rs.randn(B, N, 3, 3))
                                     .reshape(B, N, 9).transpose(0, 2, 1).astype(np.float32).copy()),
         "frame_t": torch.from_numpy(rs.randn(B, 4, N).astype(np.float32)),
         "movable_logits": torch.from_numpy(rs.rand(B, 5, N).astype(np.float32))}
labels = {"scene_score_labels": torch.from_numpy(rs.randint(0, 3, (B, N))),
          "scene_movable_labels": torch.from_numpy(rs.randint(0, 2, (B, 5, N)).astype(np.float32)),
          "best_frame_R": torch.from_numpy(np.linalg.qr(rs.randn(B, n, 3, 3))[0].reshape(B, n, 9)
                                           .transpose(0, 2, 1).astype(np.float32).copy()),
          "best_frame_t": torch.from_numpy(rs.randint(0, 4, (B, n))),
          "scene_score": torch.from_numpy(rs.rand(B, N).astype(np.float32))}
fix = {}
fix.update({"pred/" + k: v.numpy() for k, v in preds.items()})
fix.update({"label/" + k: v.numpy() for k, v in labels.items()})
for tag, ls in (("loss", 0.0), ("loss_smooth", 0.1)):
    out = PointNet2Loss(label_smoothing=ls, neg_weight=0.5)(preds, {k: v.clone() for k, v in labels.items()})
    fix.update({tag + "/" + k: v.numpy() for k, v in out.items()})
met = PointNet2Metric()(preds, {k: v.clone() for k, v in labels.items()})
fix.update({"metric/" + k: v.numpy() for k, v in met.items()})
np.savez_compressed(os.path.join(HERE, "pn2cls_loss.npz"), **fix)
print({k: (v.shape, float(np.mean(v))) for k, v in fix.items() if k.split("/")[0] in ("loss", "loss_smooth", "metric")})
