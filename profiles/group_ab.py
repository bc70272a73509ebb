"""A/B of the standalone group_points forward (reference interface: fp32 channel-first, int64 indices): plain gather vs the
shared-memory-staged variant, at the shapes of BASELINE config 3 and of the PN2_CLS module path (64 clouds).
    python profiles/group_ab.py > gpurun_out/group_ab.txt"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from s4g_release_b200._lib import lib  # noqa: E402
from s4g_release_b200.network_models.models.pointnet2_utils import pn2_ext  # noqa: E402

PEAK = 6552.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, iters=10):
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
    return sorted(ms)[len(ms) // 2]


print("B C N M K | plain ms (frac of HBM) | staged ms (frac) | speed-up     [bytes = B M K (8 + 2*4*C), SURVEY 8d]")
for B, C, N, M, K in [(64, 3, 16384, 512, 32), (64, 3, 16384, 8192, 64), (64, 3, 32768, 4096, 64), (64, 3, 25600, 5120, 64),
                      (16, 256, 5120, 1024, 64), (16, 512, 1024, 256, 64), (1, 3, 16384, 8192, 64), (8, 64, 8192, 2048, 32)]:
    g = torch.Generator().manual_seed(N)
    x = torch.randn(B, C, N, generator=g).cuda()
    idx = torch.randint(0, N, (B, M, K), generator=g).cuda()
    by = B * M * K * (8.0 + 8.0 * C)
    res = []
    for mode in (0, 1):
        lib.s4g_group_points_set_staged(mode)
        ms = timed(lambda: pn2_ext.group_points_forward(x, idx))
        res.append((ms, by / ms / 1e6 / PEAK))
    lib.s4g_group_points_set_staged(1)
    print("%d %d %d %d %d | %.4f (%.3f) | %.4f (%.3f) | %.2fx" % (B, C, N, M, K, res[0][0], res[0][1], res[1][0], res[1][1],
                                                                res[0][0] / res[1][0]))
