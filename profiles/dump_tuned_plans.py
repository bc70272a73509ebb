"""Re-time every chain shape of PN2_CLS on this GPU and print the plan table entries (to regenerate
s4g_release_b200/tuned_plans.json):   python profiles/dump_tuned_plans.py > gpurun_out/tuned_plans_run.json"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import seeded_model  # noqa: E402
from s4g_release_b200 import engine  # noqa: E402

t0 = time.time()
eng = engine.FusedPointNet2(seeded_model().cuda(), autotune="force")
torch.cuda.synchronize()
out = engine.export_tuned_plans()
out["_seconds"] = round(time.time() - t0, 2)
print(json.dumps(out, indent=1))
