"""FPS timing probe: python profiles/fps_probe.py  (env S4G_FPS_DUAL=0/1) — per-iteration time vs batch / cluster size."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from s4g_release_b200._lib import lib, ptr, stream_ptr, check
dev = torch.device("cuda")
for B, N, M in [(1, 25600, 5120), (16, 25600, 5120), (32, 25600, 5120), (64, 25600, 5120), (128, 25600, 2048), (1, 12800, 2048), (64, 12800, 2048),
                (1, 5120, 1024), (64, 5120, 1024), (32, 5120, 1024), (1, 1024, 256), (64, 1024, 256)]:
    x = torch.rand(B, 3, N, device=dev)
    idx = torch.empty(B, M, dtype=torch.int32, device=dev)
    def run():
        check(lib.s4g_farthest_point_sample_f32_i32(ptr(x), B, N, M, ptr(idx), stream_ptr(dev)), "fps")
    run(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print("cluster=%s B %3d N %5d M %4d: %7.3f ms  %.3f us/iteration" % (os.environ.get("S4G_FPS_CLUSTER", "auto"), B, N, M, ms, 1e3 * ms / (M - 1)))
