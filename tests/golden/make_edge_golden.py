"""Golden fixture for the EDGEPN2D model and the PN2 metric from the REFERENCE's own classes (build container only):
network_models/models/EdgePointNet2Down.py::EdgePointNet2Down (PN2 with EdgeSAModule levels) and
PointNet2.py::PointNet2Metric, torch-CPU with the C restatement of the operators injected as pn2_ext.
(EdgePointNet2DownUp cannot be instantiated in the reference — NameError: SharedMLP, EdgePointNet2DownUp.py:65 — so
there is nothing to record for it.)  -> tests/golden/edge_models.npz"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from make_golden import import_reference, seed_reference_weights  # noqa: E402

import_reference()
from grasp_proposal.network_models.models.EdgePointNet2Down import EdgePointNet2Down  # noqa: E402
from grasp_proposal.network_models.models.PointNet2 import PointNet2Metric  # noqa: E402

try:
    from grasp_proposal.network_models.models.EdgePointNet2DownUp import EdgePointNet2DownUp
    EdgePointNet2DownUp(score_classes=3)
    downup = "constructible"
except NameError as e:
    downup = "NameError: %s" % e
print("reference EdgePointNet2DownUp:", downup)

CFG = dict(score_classes=3, num_centroids=(256, 64, 16, 0), radius=(0.1, 0.2, 0.4, -1.0), num_neighbours=(16, 16, 8, -1),
           sa_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128, 256)),
           fp_channels=((64, 64), (64, 32), (32, 32), (32, 32, 16)), num_fp_neighbours=(0, 3, 3, 3), seg_channels=(32,),
           dropout_prob=0.5)
torch.manual_seed(0)
net = seed_reference_weights(EdgePointNet2Down(**CFG))
torch.nn.init.normal_(net.t_logit.weight, std=0.05)
net.eval()
rs = np.random.RandomState(9)
points = torch.from_numpy(rs.rand(2, 3, 1024).astype(np.float32))
with torch.no_grad():
    out = net({"scene_points": points})
n = 100
labels = {"scene_score_labels": torch.from_numpy(rs.randint(0, 3, (2, 1024))),
          "scene_movable_labels": torch.from_numpy(rs.randint(0, 2, (2, 5, 1024)).astype(np.float32)),
          "best_frame_R": torch.from_numpy(np.linalg.qr(rs.randn(2, n, 3, 3))[0].reshape(2, n, 9).transpose(0, 2, 1)
                                           .astype(np.float32).copy()),
          "best_frame_t": torch.from_numpy(rs.rand(2, 3, n).astype(np.float32)),
          "scene_score": torch.from_numpy(rs.rand(2, 1024).astype(np.float32))}
metric = PointNet2Metric()(out, labels)
fix = {"points": points.numpy(), "downup_in_reference": np.array(downup)}
fix.update({"sd/" + k: v.numpy() for k, v in net.state_dict().items()})
fix.update({"out/" + k: v.numpy() for k, v in out.items()})
fix.update({"label/" + k: v.numpy() for k, v in labels.items()})
fix.update({"metric/" + k: v.numpy() for k, v in metric.items()})
np.savez_compressed(os.path.join(HERE, "edge_models.npz"), **fix)
print({k: v.shape for k, v in fix.items() if "sd/" not in k})
