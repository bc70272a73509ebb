"""One launch of each GEMM epilogue flavour at the 10.5 M x 128 -> 128 shape, for ncu:
    ncu --set full --import-source on --clock-control none -k regex:gemm_bf16 -o gpurun_out/gemm_bwd python profiles/gemm_bwd_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from s4g_release_b200.train_engine import gemm, gemm_bwd  # noqa: E402

BF = torch.bfloat16
P, K, N = int(os.environ.get("S4G_PROBE_ROWS", 10485760 // 4)), 128, 128
a = torch.randn(P, K, device="cuda").to(BF)
b = (torch.randn(N, K, device="cuda") / K ** 0.5).to(BF)
y = torch.randn(P, N, device="cuda").to(BF)
sc = torch.rand(N, device="cuda") + 0.5
sh = torch.randn(N, device="cuda") * 0.3
gemm(a, b)
gemm(a, b, stats=True)
gemm_bwd(a, b, y, sc, sh)
torch.cuda.synchronize()
