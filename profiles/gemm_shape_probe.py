"""One plain + one stats launch of the training GEMM per shape in S4G_PROBE_SHAPES="P,K,N;P,K,N", for ncu."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from s4g_release_b200.train_engine import gemm  # noqa: E402

BF = torch.bfloat16
for spec in os.environ.get("S4G_PROBE_SHAPES", "2097152,256,512").split(";"):
    P, K, N = (int(v) for v in spec.split(","))
    a = torch.randn(P, K, device="cuda").to(BF)
    b = (torch.randn(N, K, device="cuda") / K ** 0.5).to(BF)
    gemm(a, b)
    gemm(a, b, stats=True)
    torch.cuda.synchronize()
    del a, b
