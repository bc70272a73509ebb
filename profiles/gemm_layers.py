"""Per-layer timing of the training GEMM (csrc/gemm_bf16.cu) at the shapes of one PN2_CLS step with 32 scenes:
    python profiles/gemm_layers.py > gpurun_out/gemm_layers.txt
CUDA events, 5 repetitions after 2 warm-ups (operands far larger than L2), algorithmic bytes 2 (P K + N K + P N) (+ 2 P N
for the y rows the BWD epilogue reads) against the measured copy bandwidth."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from s4g_release_b200._lib import lib  # noqa: E402
from s4g_release_b200.train_engine import gemm, gemm_bwd  # noqa: E402

BF = torch.bfloat16
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
S = int(os.environ.get("S4G_PROFILE_BATCH", "32"))
R0, R1, R2, RP = S * 5120 * 64, S * 1024 * 64, S * 256 * 64, S * 25600
# (rows, K, N, kind): forward layers (stats) and input-gradient layers (plain / bwd)
FWD = [(R0, 8, 128), (R0, 128, 128), (R0, 128, 256), (R1, 264, 256), (R1, 256, 256), (R1, 256, 512), (R2, 520, 512),
       (R2, 512, 512), (R2, 512, 1024), (RP, 256, 512), (RP, 512, 256), (RP, 256, 256), (RP, 256, 128), (RP, 384, 256)]
BWD = [(R0, 256, 128), (R0, 128, 128), (R1, 512, 256), (R1, 256, 256), (R1, 256, 256), (R2, 1024, 512), (R2, 512, 512),
       (RP, 128, 256), (RP, 256, 256), (RP, 256, 512), (RP, 512, 256)]


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    tiles = os.environ.get("S4G_GEMM_LAYERS_TILES") == "1"  # compare 128- with 256-column tiles instead of epilogue groups
    print("%-28s %10s %10s %10s %10s   (ms; fraction of %.0f GB/s)" % (("rows x K -> N", "plain n128", "plain n256", "fused n128", "fused n256", PEAK)
          if tiles else ("rows x K -> N", "plain g1", "plain g2", "fused g1", "fused g2", PEAK)))
    for kind, shapes in (("fwd+stats", FWD), ("dX+reduce", BWD)):
        for P, K, N in shapes:
            a = torch.randn(P, K, device="cuda").to(BF)
            b = (torch.randn(N, K, device="cuda") / K ** 0.5).to(BF)
            y = torch.randn(P, N, device="cuda").to(BF) if kind != "fwd+stats" else None
            sc = torch.rand(N, device="cuda") + 0.5
            sh = torch.randn(N, device="cuda") * 0.3
            bytes_plain = 2.0 * (P * K + N * K + P * N)
            bytes_fused = bytes_plain + (2.0 * P * N if y is not None else 0.0)
            cells = []
            for fused in (False, True):
                for groups in (1, 2):
                    if tiles:
                        lib.s4g_gemm_bf16_set_tile_n(128 * groups)
                    else:
                        lib.s4g_gemm_bf16_set_epilogue_groups(groups)
                    if not fused:
                        ms = timed(lambda: gemm(a, b))
                    elif y is None:
                        ms = timed(lambda: gemm(a, b, stats=True))
                    else:
                        ms = timed(lambda: gemm_bwd(a, b, y, sc, sh))
                    by = bytes_fused if fused else bytes_plain
                    cells.append("%.3f/%.2f" % (ms, by / (ms * 1e-3) / 1e9 / PEAK))
            lib.s4g_gemm_bf16_set_epilogue_groups(0)
            lib.s4g_gemm_bf16_set_tile_n(0)
            print("%-9s %9d x %4d -> %4d %10s %10s %10s %10s" % ((kind, P, K, N) + tuple(cells)))
            del a, b, y


main()
