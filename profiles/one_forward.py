"""One fused PN2_CLS forward at the BASELINE config[1] shape (64 scenes x 25 600 points) for ncu captures.

    ncu --set full --clock-control none --import-source on -k regex:'mlp_chain|fps_kernel|ball_query|three_nn' \
        -c 20 -o gpurun_out/prof python profiles/one_forward.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import seeded_model, synthetic_scenes  # noqa: E402
from s4g_release_b200.engine import FusedPointNet2  # noqa: E402

B = int(os.environ.get("S4G_PROFILE_BATCH", "64"))
net = seeded_model().cuda()
eng = FusedPointNet2(net)
base = synthetic_scenes(min(B, 8), 1000)
scenes = base.repeat((B + base.shape[0] - 1) // base.shape[0], 1, 1)[:B].contiguous().cuda()
for _ in range(int(os.environ.get("S4G_PROFILE_ITERS", "1"))):
    out = eng.forward(scenes)
torch.cuda.synchronize()
print({k: tuple(v.shape) for k, v in out.items()})
