"""One fused MLP chain of PN2_CLS at the BASELINE config[1] shape, for ncu captures:
    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:mlp_chain -c 1 -o gpurun_out/prof \
        python profiles/one_chain.py head0        (names: sa0 sa1 sa2 fp1 fp2 head0)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import seeded_model  # noqa: E402
from s4g_release_b200.engine import FusedPointNet2  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "head0"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
eng = FusedPointNet2(seeded_model().cuda())
cfg = eng.cfg
g = torch.Generator(device="cuda").manual_seed(0)
lv_n = [25600] + list(cfg["num_centroids"])
if name.startswith("sa"):
    i = int(name[2])
    ch = eng.sa_chains[i]
    Nn, M, K = lv_n[i], lv_n[i + 1], cfg["num_neighbours"][i]
    xyz = torch.rand(B, 3, Nn, device="cuda", generator=g)
    ctr = xyz[:, :, :M].contiguous()
    nbr = torch.randint(0, Nn, (B, M, K), device="cuda", dtype=torch.int32, generator=g)
    fc = ch.all_cin[0] - 3
    feat = torch.randn(B * Nn, fc, device="cuda", generator=g).to(torch.bfloat16) if fc else None
    run = lambda: ch.run_gather(feat, xyz, ctr, nbr)
elif name.startswith("fp"):
    i = int(name[2])
    ch = eng.fp_chains[i][0]
    x = torch.randn(B * lv_n[-2 - i], ch.cin[0], device="cuda", generator=g).to(torch.bfloat16)
    run = lambda: ch.run_rows(x)
else:
    ch = eng.head_chains[int(name[4])]
    x = torch.randn(B * lv_n[0], 256, device="cuda", generator=g).to(torch.bfloat16)
    run = lambda: ch.run_rows(x, n_points=lv_n[0])
for _ in range(3):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()  # ncu --profile-from-start off: skip the autotuner's candidate launches
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(name, ch.info())
