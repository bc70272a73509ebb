"""One fused PN2_CLS forward at the BASELINE config[1] shape (64 scenes x 25 600 points) for ncu captures.

    ncu --profile-from-start off --metrics gpu__time_duration.sum,... --clock-control none --csv \
        --log-file gpurun_out/launches.csv python profiles/one_forward.py          (see profiles/ncu_to_json.py)
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
# the engine's autotuner launches hundreds of candidate chains while it is built: profile only this forward
# (ncu --profile-from-start off)
torch.cuda.profiler.start()
out = eng.forward(scenes)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print({k: tuple(v.shape) for k, v in out.items()})
