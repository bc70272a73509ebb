import os, sys, torch
sys.path.insert(0, os.getcwd())
from bench import seeded_model
from s4g_release_b200.engine import FusedPointNet2
B=64
eng = FusedPointNet2(seeded_model().cuda())
cfg = eng.cfg
g = torch.Generator(device="cuda").manual_seed(0)
lv_n = [25600] + list(cfg["num_centroids"])
def timed(fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts=[]
    for _ in range(5):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
out=[]
for i, chain in enumerate(eng.sa_chains):
    Nn, M, K = lv_n[i], lv_n[i + 1], cfg["num_neighbours"][i]
    xyz = torch.rand(B, 3, Nn, device="cuda", generator=g); ctr = xyz[:, :, :M].contiguous()
    nbr = torch.randint(0, Nn, (B, M, K), device="cuda", dtype=torch.int32, generator=g)
    fc = chain.all_cin[0] - 3
    feat = torch.randn(B * Nn, fc, device="cuda", generator=g).to(torch.bfloat16) if fc else None
    out.append(("sa%d"%i, timed(lambda: chain.run_gather(feat, xyz, ctr, nbr))))
for i, chains in enumerate(eng.fp_chains):
    rows = B * lv_n[-2 - i]
    for k, ch in enumerate(chains):
        x = torch.randn(rows, ch.cin[0], device="cuda", generator=g).to(torch.bfloat16)
        out.append(("fp%d.%d"%(i,k), timed(lambda: ch.run_rows(x))))
x = torch.randn(B*25600, 256, device="cuda", generator=g).to(torch.bfloat16)
out.append(("head0", timed(lambda: eng.head_chains[0].run_rows(x, n_points=25600))))
print(os.environ.get("S4G_LIB_PATH","new"), " ".join("%s %.3f"%(n,t) for n,t in out))
