"""Per-role cycle breakdown of every fused MLP chain of PN2_CLS at the BASELINE config[1] shape, from the
kernel's optional counters (s4g_chain_set_profile).  Each chain runs alone on synthetic inputs of the
right shape, timed with CUDA events.   python profiles/chain_prof.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import seeded_model  # noqa: E402
from s4g_release_b200.engine import FusedPointNet2  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
TRACE = sys.argv[2].split(",") if len(sys.argv) > 2 else []
net = seeded_model().cuda()
eng = FusedPointNet2(net)
if os.environ.get("S4G_PROF_PLAN"):  # "slots,pairs,coop,subs": rebuild the set-abstraction chains under these constraints
    from s4g_release_b200.chain import IN_GATHER, OUT_MAXPOOL, MlpChain
    sl, pr, co, su = [int(v) for v in os.environ["S4G_PROF_PLAN"].split(",")]
    fc = 0
    for i, layers in enumerate(eng.sa):
        try:
            eng.sa_chains[i] = MlpChain([(w, b, True) for w, b in layers], "cuda", IN_GATHER, fc, OUT_MAXPOOL,
                                        group=eng.cfg["num_neighbours"][i], slots=sl, pairs=pr, coop=co, subs=su)
        except RuntimeError:
            pass
        fc = layers[-1][0].shape[0]
cfg = eng.cfg
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
NAMES = ["prod_total", "prod_wait_stage", "mma_total", "mma_wait_act", "mma_wait_tmem", "mma_wait_w", "epi_total",
         "epi_wait_acc", "ld_wait_slot", "ld_wait_cp", "epi_work", "ld_total", "epi_wait_slot"]


def timed(name, chain, fn, rows):
    cnt = torch.zeros(148 * 16 + 4 * 128 + 4 * 32 + 2 * 4 * 64, dtype=torch.int64, device=dev)
    chain.set_profile(cnt)
    fn()
    torch.cuda.synchronize()
    cnt.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    chain.set_profile(None)
    c = cnt[:148 * 16].reshape(148, 16).double().mean(0).cpu().tolist()
    trace = cnt[148 * 16:].cpu().tolist()
    info = chain.info()
    tiles = (rows + 127) // 128 / 148.0
    tf = chain.flops(rows) / ms / 1e9
    print("%-10s %.3f ms  %.0f TFLOP/s  tiles/SM %.1f  cyc/tile %.0f (plan %d, mma %d)  slots %d stages %d jobs %d" %
          (name, ms, tf, tiles, c[2] / tiles, info["sim_cycles"], info["mma_cycles"], info["slots"], info["stages"],
           info["n_jobs"]))
    print("           " + "  ".join("%s %.0f" % (n, v / tiles) for n, v in zip(NAMES, c)))
    if TRACE and name in TRACE:
        nj, ne = info["n_jobs"], 32
        t0 = min(v for v in trace if v > 0)
        print("   MMA jobs of CTA 0, 3rd tile (cycles from the tile's first event): wait_begin waits_done issued probed")
        for j in range(nj):
            r = trace[4 * j: 4 * j + 4]
            print("     job %2d: %6d %6d %6d %6d   (wait %5d, issue %5d, probe %5d)" % (j, r[0] - t0, r[1] - t0, r[2] - t0, r[3] - t0,
                  r[1] - r[0], r[2] - r[1], r[3] - r[2]))
        print("   epilogue jobs (first warp of the group that ran it): wait_begin acc_full done")
        for j in range(ne):
            r = trace[4 * 128 + 4 * j: 4 * 128 + 4 * j + 3]
            if r[0] > 0:
                print("     epi %2d: %6d %6d %6d   (waited %5d, work %5d)" % (j, r[0] - t0, r[1] - t0, r[2] - t0, r[1] - r[0], r[2] - r[1]))
        print("   loader jobs of the 3rd and 4th tile (same clock): begin slot_free issued")
        for k in range(2):
            for j in range(64):
                r = trace[4 * 128 + 4 * 32 + 4 * (64 * k + j): 4 * 128 + 4 * 32 + 4 * (64 * k + j) + 3]
                if r[0] > 0:
                    print("     tile %d load %2d: %6d %6d %6d   (waited %5d for the slot, issue %5d)" %
                          (3 + k, j, r[0] - t0, r[1] - t0, r[2] - t0, r[1] - r[0], r[2] - r[1]))


N = cfg["num_points"] if "num_points" in cfg else 25600
lv_n = [25600] + list(cfg["num_centroids"])
feat = None
for i, chain in enumerate(eng.sa_chains):
    Nn, M, K = lv_n[i], lv_n[i + 1], cfg["num_neighbours"][i]
    xyz = torch.rand(B, 3, Nn, device=dev, generator=g)
    ctr = xyz[:, :, :M].contiguous()
    nbr = torch.randint(0, Nn, (B, M, K), device=dev, dtype=torch.int32, generator=g)
    fc = chain.all_cin[0] - 3
    feat = torch.randn(B * Nn, fc, device=dev, generator=g).to(torch.bfloat16) if fc else None
    timed("sa%d" % i, chain, lambda: chain.run_gather(feat, xyz, ctr, nbr), B * M * K)
for i, chains in enumerate(eng.fp_chains):
    rows = B * lv_n[-2 - i]
    for k, ch in enumerate(chains):
        x = torch.randn(rows, ch.cin[0], device=dev, generator=g).to(torch.bfloat16)
        timed("fp%d.%d" % (i, k), ch, lambda: ch.run_rows(x), rows)
rows = B * lv_n[0]
x = torch.randn(rows, 256, device=dev, generator=g).to(torch.bfloat16)
for k, ch in enumerate(eng.head_chains[:2]):
    timed("head%d" % k, ch, lambda: ch.run_rows(x, n_points=lv_n[0]), rows)
