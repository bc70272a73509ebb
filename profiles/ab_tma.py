"""A/B: the row chains of PN2_CLS (heads, fp) with cp.async input staging vs TMA tensor-copy input, same plans.
    python profiles/ab_tma.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from s4g_release_b200.chain import OUT_LOGITS, OUT_ROWS, MlpChain  # noqa: E402


def layers(dims, relu_last=True):
    g = torch.Generator().manual_seed(0)
    return [(torch.randn(dims[i + 1], dims[i], generator=g) / dims[i] ** 0.5, torch.zeros(dims[i + 1]),
             relu_last or i + 2 < len(dims)) for i in range(len(dims) - 1)]


def time_chain(ch, x, n_points):
    for _ in range(2):
        ch.run_rows(x, n_points=n_points)
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ch.run_rows(x, n_points=n_points)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


P = 64 * 25600
cases = [("head (256-512-256-256-128-3)", [256, 512, 256, 256, 128, 3], OUT_LOGITS, P),
         ("fp2-like (384-256-256)", [384, 256, 256], OUT_ROWS, 64 * 5120),
         ("fp1-like (768-512-512)", [768, 512, 512], OUT_ROWS, 64 * 1024),
         ("rows 128-128-128", [128, 128, 128], OUT_ROWS, P),
         ("rows 64-64-64", [64, 64, 64], OUT_ROWS, P)]
for name, dims, out_mode, rows in cases:
    L = layers(dims, relu_last=out_mode != OUT_LOGITS)
    x = torch.randn(rows, dims[0], device="cuda").to(torch.bfloat16)
    n_points = 25600 if out_mode == OUT_LOGITS else 0
    for slots in (0, 3, 4, 5):
        for subs in (1, 2):
            row = []
            for tma in (0, 1):
                try:
                    ch = MlpChain(L, "cuda", out_mode=out_mode, slots=slots, subs=subs, tma_in=tma)
                except RuntimeError:
                    row.append(None)
                    continue
                row.append(time_chain(ch, x, n_points))
            if row[0] is None and row[1] is None:
                continue
            print("%-30s slots %d subs %d  cp.async %s ms   tma %s ms" %
                  (name, slots, subs, *["%.3f" % t if t else "  -  " for t in row]), flush=True)


# gathered set-abstraction chains (levels 1 and 2 of PN2_CLS): cp.async gather vs tile::gather4
from s4g_release_b200.chain import IN_GATHER, OUT_MAXPOOL  # noqa: E402
for name, feat_c, dims, N, M, K in [("sa1 (259-256-256-512)", 256, [256, 256, 512], 5120, 1024, 64),
                                    ("sa2 (515-512-512-1024)", 512, [512, 512, 1024], 1024, 256, 64)]:
    B = 64
    L = layers([feat_c + 3] + dims)
    g = torch.Generator(device="cuda").manual_seed(0)
    xyz = torch.rand(B, 3, N, device="cuda", generator=g)
    ctr = xyz[:, :, :M].contiguous()
    near = torch.randint(0, 256, (B, M, K), device="cuda", generator=g)
    nbr = ((torch.arange(M, device="cuda").view(1, M, 1) * (N // M) + near) % N).to(torch.int32)
    feat = torch.randn(B * N, feat_c, device="cuda", generator=g).to(torch.bfloat16)

    class G:
        def __init__(self, ch):
            self.ch = ch

        def run_rows(self, x, n_points=0):
            return self.ch.run_gather(feat, xyz, ctr, nbr)

    for slots in (0, 3, 4, 5):
        for pairs in (-1, 0):
            row = []
            for tma in (0, 1):
                try:
                    ch = MlpChain(L, "cuda", IN_GATHER, feat_c, OUT_MAXPOOL, group=K, slots=slots, pairs=pairs, tma_in=tma)
                except RuntimeError:
                    row.append(None)
                    continue
                row.append(time_chain(G(ch), None, 0))
            if row[0] is None and row[1] is None:
                continue
            print("%-30s slots %d pairs %2d  cp.async %s ms   gather4 %s ms" %
                  (name, slots, pairs, *["%.3f" % t if t else "  -  " for t in row]), flush=True)
