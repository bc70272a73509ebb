"""BASELINE config 3 — micro-benchmark sweep of the geometry kernels:
FPS + ball query + group_points for N in {16 384 … 262 144} x npoint in {512 … 8192} x nsample in {32, 64},
B in {1, 64}; points U[0,1)^3 from RandomState(N), centroids = FPS output, radius r = (3 nsample / (4 pi N))^(1/3).

Each kernel is timed alone with CUDA events (>= 3 warm-ups, L2 flushed between iterations) through the
reference-shaped `pn2_ext` operators (int64 indices, channel-first tensors) and reported in microseconds per cloud
and as a fraction of the measured HBM peak on the ALGORITHMIC bytes of §8(d):
    FPS          B (M-1) N 16   (what the reference streams; > 100 % is the signature of on-chip residency)
    ball query   B (M N 12 + M K 8 + M 8)
    group_points B M K (8 + 2 * 4 * C), C = 3
One JSON object per line on stdout; `--out FILE` also writes a markdown table.
    python bench_micro.py [--quick] [--out profiles/r01/micro_sweep.md]
"""
import argparse
import json
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def timed(fn, flush, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    return float(np.median(ms))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="corners of the sweep only")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    from s4g_release_b200.network_models.models.pointnet2_utils import pn2_ext
    peak, src = hbm_peak()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    Ns = [16384, 32768, 65536, 131072, 262144]
    Ms = [512, 1024, 2048, 4096, 8192]
    Ks = [32, 64]
    Bs = [1, 64]
    if args.quick:
        Ns, Ms, Ks = [16384, 262144], [512, 8192], [64]
    rows = []
    for B in Bs:
        for N in Ns:
            pts = torch.from_numpy(np.random.RandomState(N).rand(B, 3, N).astype(np.float32)).cuda()
            for M in Ms:
                iters = 3 if B * N * M > 2 ** 34 else 10
                idx = pn2_ext.farthest_point_sample(pts, M)
                t_fps = timed(lambda: pn2_ext.farthest_point_sample(pts, M), flush, iters)
                ctr = torch.gather(pts, 2, idx.unsqueeze(1).expand(-1, 3, -1)).contiguous()
                for K in Ks:
                    r = (3.0 * K / (4.0 * math.pi * N)) ** (1.0 / 3.0)
                    nbr, cnt = pn2_ext.ball_query(pts, ctr, r, K)
                    t_bq = timed(lambda: pn2_ext.ball_query(pts, ctr, r, K), flush, iters)
                    t_gp = timed(lambda: pn2_ext.group_points_forward(pts, nbr), flush, iters)
                    b_fps = B * (M - 1) * N * 16.0
                    b_bq = B * (M * N * 12.0 + M * K * 8.0 + M * 8.0)
                    b_gp = B * M * K * (8.0 + 2 * 4 * 3)
                    row = {"B": B, "N": N, "npoint": M, "nsample": K, "radius": round(r, 5),
                           "mean_neighbours": round(float(cnt.float().mean()), 2),
                           "fps_us_per_cloud": round(t_fps * 1e3 / B, 2), "fps_frac_hbm": round(b_fps / t_fps / 1e6 / peak, 3),
                           "ball_query_us_per_cloud": round(t_bq * 1e3 / B, 2),
                           "ball_query_frac_hbm": round(b_bq / t_bq / 1e6 / peak, 3),
                           "group_us_per_cloud": round(t_gp * 1e3 / B, 2), "group_frac_hbm": round(b_gp / t_gp / 1e6 / peak, 4),
                           "hbm_peak_gbs": peak, "peak_source": src}
                    rows.append(row)
                    print(json.dumps(row), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            f.write("# Geometry micro-benchmark sweep (BASELINE config 3), one B200, HBM peak %.0f GB/s (%s)\n\n" % (peak, src))
            f.write("us = microseconds per cloud; frac = algorithmic bytes (SURVEY §8d) / time / HBM peak\n\n")
            f.write("| B | N | npoint | nsample | mean nbrs | FPS us | FPS frac | BQ us | BQ frac | group us | group frac |\n")
            f.write("|---|---|---|---|---|---|---|---|---|---|---|\n")
            for r in rows:
                f.write("| %d | %d | %d | %d | %.1f | %.1f | %.2f | %.1f | %.2f | %.1f | %.3f |\n" % (
                    r["B"], r["N"], r["npoint"], r["nsample"], r["mean_neighbours"], r["fps_us_per_cloud"], r["fps_frac_hbm"],
                    r["ball_query_us_per_cloud"], r["ball_query_frac_hbm"], r["group_us_per_cloud"], r["group_frac_hbm"]))


if __name__ == "__main__":
    main()
