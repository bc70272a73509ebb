#!/usr/bin/env python
"""BASELINE config 4: PN2_CLS TRAINING step — 32 scenes x 25 600 points per GPU, PointNet2Loss, Adam (lr 1e-3), one
process per GPU, ONE flat NCCL all-reduce of the 6.63 M fp32 gradients per step (s4g_release_b200/train.py).

    python bench_train.py [--gpus N] [--steps K] [--warmup W] [--batch 32]      (N > 1: under torchrun like bench.py)

`--path fused` (default): s4g_release_b200/train_engine.py — forward, loss and the hand-written backward on the
B200-native training kernels: tcgen05 GEMMs over channel-last bf16 rows (csrc/gemm_bf16.cu), fused BatchNorm-statistics /
normalise / ReLU / dropout / max-pool passes and their backward (csrc/train_ops.cu), the sm_100a geometry operators.
`--path module`: the reference-shaped nn.Module stack on the seven sm_100a operators under torch autograd; the shared
MLPs are then torch convolutions (fp32 storage; TF32 math by default, like the unmodified reference on this GPU) — the
round-1 path, kept as the comparison.  `value` = scenes/s with inputs and
labels resident; `e2e` adds the pinned-host -> device copy of the clouds and labels and the device -> host read of
the loss every step.  Dropout enabled (timing run), per-replica BatchNorm statistics, weak scaling.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import ClockSampler, NUM_POINTS, synthetic_scenes  # noqa: E402

METRIC = "S4G scenes/sec (PN2_CLS training step, 25600 points/scene)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="scenes per GPU per step")
    ap.add_argument("--num-frame", type=int, default=4000)
    ap.add_argument("--path", default="fused", choices=["fused", "module"])
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp32"],
                    help="tf32 = torch's default for cuDNN convolutions, i.e. what the unmodified reference runs on this "
                         "GPU (SURVEY.md §8a6); fp32 = IEEE")
    ap.add_argument("--fused-bwd-reduce", action="store_true", help="A/B: BatchNorm-backward sums in the dX GEMM's epilogue")
    ap.add_argument("--fp-split", action="store_true", help="A/B: finest propagation level with its conv on the sparse rows")
    ap.add_argument("--interp-bwd-scatter", action="store_true", help="A/B: interpolation gradient by atomic scatter")
    ap.add_argument("--no-sparse-pool-reduce", action="store_true", help="A/B: pooled blocks reduce over all G*K rows")
    ap.add_argument("--epilogue-groups", type=int, default=0, choices=[0, 1, 2], help="A/B: GEMM epilogue warp groups")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    from s4g_release_b200 import _lib
    from s4g_release_b200.network_models.models.PointNet2_tcls import PN2_CLS_CONFIG, PointNet2, PointNet2Loss
    from s4g_release_b200.train import Trainer, synthetic_labels

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench_train.py needs a CUDA device (there is no CPU path for the product)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cudnn.allow_tf32 = args.precision == "tf32"
    torch.backends.cuda.matmul.allow_tf32 = args.precision == "tf32"
    torch.backends.cudnn.benchmark = True

    from s4g_release_b200 import train_engine
    train_engine.FUSED_BWD_REDUCE = args.fused_bwd_reduce
    train_engine.SPARSE_POOL_REDUCE = not args.no_sparse_pool_reduce
    train_engine.FP_LINEAR_SPLIT = args.fp_split
    train_engine.INTERP_BWD_GATHER = not args.interp_bwd_scatter
    if args.epilogue_groups:
        _lib.lib.s4g_gemm_bf16_set_epilogue_groups(args.epilogue_groups)
    B = args.batch
    torch.manual_seed(0)
    model = PointNet2(**PN2_CLS_CONFIG).to(dev)
    trainer = Trainer(model, PointNet2Loss(), fused=(args.path == "fused"))
    host_x = synthetic_scenes(B, 1000 + rank * B).pin_memory()
    host_y = {k: v.pin_memory() for k, v in synthetic_labels(B, NUM_POINTS, args.num_frame, 2000 + rank * B).items()}
    x = host_x.to(dev)
    y = {k: v.to(dev) for k, v in host_y.items()}

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    clocks.__enter__()
    for _ in range(args.warmup):
        trainer.step({"scene_points": x}, y)
    sync_all()
    launches0 = _lib.lib.s4g_launch_count()
    t_w0 = time.time()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        losses = trainer.step({"scene_points": x}, y)
    b.record()
    sync_all()
    clocks.window(t_w0, time.time())
    ms_local = a.elapsed_time(b) / args.steps
    launches = (_lib.lib.s4g_launch_count() - launches0) // args.steps
    peak_mem = torch.cuda.max_memory_allocated(dev)

    # end to end: host buffers in, loss out, every step
    e2e = []
    h2d = host_x.numel() * 4 + sum(v.numel() * v.element_size() for v in host_y.values())
    for i in range(1 + args.steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        xs = host_x.to(dev, non_blocking=True)
        ys = {k: v.to(dev, non_blocking=True) for k, v in host_y.items()}
        losses = trainer.step({"scene_points": xs}, ys)
        total = float(sum(losses.values()))  # device -> host read of the loss
        if i >= 1:
            e2e.append(1e3 * (time.perf_counter() - t0))
            clocks.window(time.time() - e2e[-1] * 1e-3, time.time())
    e2e_local = sum(e2e) / len(e2e)
    time.sleep(0.05)
    clocks.__exit__(None, None, None)

    tt = torch.tensor([ms_local, e2e_local], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms, e2e_ms = tt[0].item(), tt[1].item()
    if rank == 0:
        line = {
            "metric": METRIC, "value": world * B / (ms * 1e-3), "unit": "scenes/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.path == "fused" else args.precision, "data": "synthetic",
            "config": {"workload": "PN2_CLS training step (BASELINE config[3]): forward + PointNet2Loss + backward + Adam, "
                                   "synthetic tabletop clouds, %d points/scene, %d labelled frames" % (NUM_POINTS, args.num_frame),
                       "scenes_per_gpu_per_step": B,
                       "parallelism": "dp%d, one flat all-reduce of 6.63 M fp32 gradients per step" % world,
                       "path": ("fused training kernels: tcgen05 bf16 GEMMs + fused BN / ReLU / dropout / max-pool passes, "
                                "hand-written backward (train_engine.py)") if args.path == "fused" else
                               ("module path: sm_100a pn2_ext operators + torch convolutions (%s), torch autograd" %
                                ("TF32 tensor cores, fp32 storage: torch's default, as the reference would run"
                                 if args.precision == "tf32" else "IEEE fp32")),
                       "l2": "activations (tens of GB per step) exceed L2; no explicit flush"},
            "clocks": clocks.summary(),
            "e2e": {"value": world * B / (e2e_ms * 1e-3), "unit": "scenes/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * len(losses)},
            "gpu_launches": int(launches) * args.steps, "gpu_launches_per_step": int(launches),
            "loss": total, "peak_memory_GB": round(peak_mem / 2**30, 2),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
