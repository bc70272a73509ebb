"""Prints the slot counts the engine's autotuner pins for every chain shape of PN2_CLS and how long tuning took."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from s4g_release_b200 import engine
t0 = time.time()
eng = engine.FusedPointNet2(bench.seeded_model().cuda())
torch.cuda.synchronize()
print("engine build + autotune: %.2f s" % (time.time() - t0))
for k, v in engine._TUNED_SLOTS.items():
    print(" ", [c for c in k[0]], "in_mode", k[1], "feat_c", k[2], "out_mode", k[3], "-> (slots, pairs, coop, subs) =", v)
