"""Run ONE chain of PN2_CLS under pinned planner constraints (debugging / tuning aid):
    python profiles/try_plan.py sa1 <slots> <pairs> <coop> <subs>"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from s4g_release_b200 import engine
from s4g_release_b200.chain import IN_GATHER, IN_ROWS, OUT_LOGITS, OUT_MAXPOOL, OUT_ROWS, MlpChain
name, slots, pairs, coop, subs = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
net = bench.seeded_model().cuda()
eng = engine.FusedPointNet2(net, autotune=False)
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
i = int(name[2])
layers = [(w, b, True) for w, b in eng.sa[i]]
feat_c = 0 if i == 0 else eng.sa[i - 1][-1][0].shape[0]
K = eng.cfg["num_neighbours"][i]
try:
    ch = MlpChain(layers, dev, IN_GATHER, feat_c, OUT_MAXPOOL, group=K, slots=slots, pairs=pairs, coop=coop, subs=subs)
except RuntimeError as e:
    print(name, sys.argv[2:], "no plan"); sys.exit(0)
rows = 148 * 48 * 128
M = rows // K; N = 4 * M
xyz = torch.rand(1, 3, N, device=dev, generator=g); ctr = xyz[:, :, :M].contiguous()
near = torch.randint(0, 512, (1, M, K), device=dev, dtype=torch.int64, generator=g)
nbr = ((torch.arange(M, device=dev).view(1, M, 1) * (N // M) + near) % N).to(torch.int32)
feat = torch.randn(N, feat_c, device=dev, generator=g).to(torch.bfloat16) if feat_c else None
ch.run_gather(feat, xyz, ctr, nbr); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); ch.run_gather(feat, xyz, ctr, nbr); b.record(); torch.cuda.synchronize()
print(name, sys.argv[2:], ch.info()["slots"], ch.info()["stages"], "%.3f ms" % a.elapsed_time(b))
