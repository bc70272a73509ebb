#!/usr/bin/env python
"""Headline benchmark: S4G (PN2_CLS) inference scenes/s on B200 — BASELINE.json config[1]:
batch 64 synthetic single-view tabletop clouds at the reference's num_points (25 600) per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path (the whole PN2_CLS forward: FPS, ball query, fused SA / FP / head
MLP chains) over one batch.  `value` = scenes/s with the inputs already resident in HBM, timed with
CUDA events on the launching stream (per-step event pairs, L2 flushed between steps, max over ranks).
`e2e` = the same metric through the public module API with HOST buffers: pinned host -> device copy of
the clouds, forward, device -> pinned host copy of the four prediction tensors, every step (the public call takes
the pinned result buffers — `model(batch, host_out=...)` — so each head's copy overlaps the next head's compute).
Weak scaling: each rank owns its own 64 scenes, no collective on the data path.

`--impl reference` times the reference model on the box's HOST cores: the oracle restatement of the
reference's python modules on the C restatement of its CUDA-only ops (the reference has no CPU
implementation of those), all host threads, each step a bounded sample (1 scene) of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "S4G scenes/sec (PN2_CLS inference, 25600 points/scene)"
UNIT = "scenes/s"
NUM_POINTS = 25600
L2_FLUSH_BYTES = 256 << 20


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "bf16_tflops_burst": d["bf16_tflops"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "bf16_tflops_burst": 1590.0,
            "source": "fallback (B200_PROFILING.md)"}


def seeded_model():
    """default init under manual_seed(0), BatchNorm statistics from Generator(1)."""
    from s4g_release_b200.network_models.models.PointNet2_tcls import PN2_CLS_CONFIG, PointNet2
    torch.manual_seed(0)
    net = PointNet2(**PN2_CLS_CONFIG)
    g = torch.Generator().manual_seed(1)
    for m in net.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            n = m.num_features
            m.weight.data = torch.rand(n, generator=g) + 0.5
            m.bias.data = torch.randn(n, generator=g) * 0.1
            m.running_mean.data = torch.randn(n, generator=g) * 0.1
            m.running_var.data = torch.rand(n, generator=g) + 0.5
    return net.eval()


def synthetic_scenes(batch, first_seed):
    from tests.inputs import tabletop_scene
    return torch.from_numpy(np.stack([tabletop_scene(first_seed + i, NUM_POINTS) for i in range(batch)]))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,timestamp")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.windows = []  # (t0, t1) wall-clock windows of the timed regions

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def __enter__(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None
        return self

    def __exit__(self, *exc):
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.p.kill()
        return False

    def summary(self):
        import datetime
        sm, mx, reasons = [], [], set()
        all_sm, all_mx = [], []
        try:
            self.f.flush()
            for line in open(self.f.name):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 10:
                    continue
                all_sm.append(float(c[1]))
                all_mx.append(float(c[2]))
                try:
                    ts = datetime.datetime.strptime(c[9], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                except ValueError:
                    ts = None
                if self.windows and ts is not None and not any(a - 0.02 <= ts <= b + 0.02 for a, b in self.windows):
                    continue
                sm.append(float(c[1]))
                mx.append(float(c[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.f.name)
        except (OSError, ValueError):
            pass
        if not sm and all_sm:  # no sample fell inside a window: report the whole run, say so
            return {"sm_mhz": statistics.median(all_sm), "sm_max_mhz": max(all_mx), "reasons": sorted(reasons),
                    "samples": len(all_sm), "note": "no sample inside the timed windows; median over the whole run"}
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def stage_model(batch):
    """ALGORITHMIC bytes / flops per launch for every stage of one step."""
    from s4g_release_b200.network_models.models.PointNet2_tcls import PN2_CLS_CONFIG as C
    B = batch
    n = [NUM_POINTS] + list(C["num_centroids"])
    K = C["num_neighbours"]
    st = {}
    cin = 0
    for i in range(3):
        N, M = n[i], n[i + 1]
        st["sa%d.fps" % i] = ("hbm", B * (M - 1) * N * 16.0)
        st["sa%d.ball_query" % i] = ("hbm", B * (M * N * 12.0 + M * K[i] * 8.0 + M * 8.0))
        st["sa%d.gather_xyz" % i] = ("hbm", B * M * (8.0 + 2 * 4 * 3))
        dims = [cin + 3] + list(C["sa_channels"][i])
        st["sa%d.mlp" % i] = ("tensor", 2.0 * B * M * K[i] * sum(a * b for a, b in zip(dims[:-1], dims[1:])))
        cin = dims[-1]
    skip = [0] + [c[-1] for c in C["sa_channels"]]
    c = skip[-1]
    for i in range(3):
        Nq, Nk = n[-2 - i], n[-1 - i]
        st["fp%d.three_nn" % i] = ("hbm", B * (Nq * Nk * 12.0 + Nq * 3 * 12.0))
        c_in = c + skip[-2 - i]
        st["fp%d.interp_concat" % i] = ("hbm", B * Nq * (3 * 12.0 + 3 * 4.0 * c + 4.0 * c))
        dims = [c_in] + list(C["fp_channels"][i])
        st["fp%d.mlp" % i] = ("tensor", 2.0 * B * Nq * sum(a * b for a, b in zip(dims[:-1], dims[1:])))
        c = dims[-1]
    seg = [c] + list(C["seg_channels"])
    per_head = sum(a * b for a, b in zip(seg[:-1], seg[1:]))
    outs = (C["score_classes"], 9, 4, C["num_removal_directions"])
    st["heads.mlp"] = ("tensor", 2.0 * B * NUM_POINTS * sum(per_head + seg[-1] * o for o in outs))
    return st


def run_reference(args, rank, world):
    """Reference arm: the reference model on the host CPU (oracle port), rank 0 only."""
    if rank != 0:
        return
    from oracle import model_cpu, pn2_ext_cpu
    torch.set_num_threads(os.cpu_count())
    net = seeded_model()
    sd = net.state_dict()
    scenes = synthetic_scenes(1, 1000)
    times = []
    with torch.no_grad():
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            model_cpu.pointnet2_forward(scenes, sd, model_cpu.PN2_CLS_CONFIG)
            dt = time.perf_counter() - t0
            if i >= args.warmup:
                times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    value = 1e3 / ms
    cores = max(pn2_ext_cpu.num_threads(), torch.get_num_threads())
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "PN2_CLS inference, synthetic tabletop clouds, 25600 points/scene (BASELINE config[1])",
                   "sample_per_step": "1 scene of the 64-scene batch"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "1 scene per step (reference python modules restated in oracle/model_cpu.py on the "
                                   "C restatement of the CUDA-only pn2_ext ops, torch-CPU fp32 conv/BN)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


TRAFFIC_FILE = "profiles/r01/ncu_launches_v21.json"


def stage_traffic():
    """DRAM bytes (read + write) per step of every stage, from the committed ncu launch list of one forward at this
    workload (dram__bytes_read.sum + dram__bytes_write.sum per launch, summed over the stage's launches)."""
    try:
        launches = json.load(open(os.path.join(ROOT, TRAFFIC_FILE)))["launches"]
    except Exception:
        return {}
    byts = lambda l: (l["dram_read_MB"] + l["dram_write_MB"]) * 1e6
    fps = [l for l in launches if "fps_kernel" in l["kernel"]]
    chains = [l for l in launches if "mlp_chain" in l["kernel"]]
    interp = [l for l in launches if "interp_concat" in l["kernel"]]
    out = {}
    if len(fps) == 3 and len(chains) == 11 and len(interp) == 3:
        for i in range(3):
            out["sa%d.fps" % i] = byts(fps[i])
            out["sa%d.mlp" % i] = byts(chains[i])
            out["fp%d.interp_concat" % i] = byts(interp[i])
        out["fp0.mlp"] = byts(chains[3]) + byts(chains[4])
        out["fp1.mlp"] = byts(chains[5])
        out["fp2.mlp"] = byts(chains[6])
        out["heads.mlp"] = sum(byts(c) for c in chains[7:11])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64, help="scenes per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mlp-backend", default="tcgen05", choices=["tcgen05", "torch"])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist
    from s4g_release_b200 import _lib
    from s4g_release_b200.engine import FusedPointNet2, StageTimer

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path for the product)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch
    net = seeded_model().to(dev)
    eng = FusedPointNet2(net, mlp_backend=args.mlp_backend)
    net.attach_engine(eng)
    host_scenes = synthetic_scenes(B, 1000 + rank * B).pin_memory()
    scenes = host_scenes.to(dev)
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- device-resident throughput ----------------
    clocks = ClockSampler(local_rank)
    clocks.__enter__()
    for _ in range(args.warmup):
        eng.forward(scenes)
    sync_all()
    launches0 = _lib.lib.s4g_launch_count()
    timers, step_events = [], []
    if True:
        sync_all()
        t_w0 = time.time()
        for _ in range(args.steps):
            flush.zero_()  # evict L2 between steps (outside the per-step event pair)
            t = StageTimer()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            eng.forward(scenes, timer=t)
            b.record()
            timers.append(t)
            step_events.append((a, b))
        sync_all()
        clocks.window(t_w0, time.time())
    launches = (_lib.lib.s4g_launch_count() - launches0) // args.steps  # kernels launched by libs4g_b200.so
    step_ms = [a.elapsed_time(b) for a, b in step_events]
    ms_local = sum(step_ms) / len(step_ms)

    # ---------------- end to end through the public API with host buffers ----------------
    out_host = None
    e2e_ms = []
    h2d = host_scenes.numel() * 4
    d2h = 0
    for i in range(args.warmup + args.steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x = host_scenes.to(dev, non_blocking=True)
        with torch.no_grad():
            if out_host is None:  # first (warm-up) step: learn the output shapes
                preds = net({"scene_points": x})
                out_host = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in preds.items()}
                d2h = sum(v.numel() * v.element_size() for v in preds.values())
                for k, v in preds.items():
                    out_host[k].copy_(v, non_blocking=True)
            else:  # the public call with pinned result buffers: every head's D2H copy overlaps the next head
                preds = net({"scene_points": x}, host_out=out_host)
        torch.cuda.synchronize()
        if i >= args.warmup:
            e2e_ms.append(1e3 * (time.perf_counter() - t0))
            clocks.window(time.time() - e2e_ms[-1] * 1e-3, time.time())
    e2e_local = sum(e2e_ms) / len(e2e_ms)
    time.sleep(0.05)
    clocks.__exit__(None, None, None)

    # ---------------- reduce over ranks (max time) ----------------
    if world > 1:
        tt = torch.tensor([ms_local, e2e_local], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, e2e = tt[0].item(), tt[1].item()
    else:
        ms, e2e = ms_local, e2e_local
    value = world * B / (ms * 1e-3)
    e2e_value = world * B / (e2e * 1e-3)

    if rank == 0:
        peaks = load_peaks()
        model = stage_model(B)
        per_stage = {}
        for t in timers:
            for name, vals in t.totals_ms().items():
                per_stage.setdefault(name, []).extend(vals)
        kernels = {}
        for name, vals in per_stage.items():
            avg = sum(vals) / len(vals)
            bound, work = model[name]
            if bound == "hbm":
                ach, peak, unit = work / (avg * 1e-3) / 1e9, peaks["hbm_gbs"], "GB/s"
            else:
                ach, peak, unit = work / (avg * 1e-3) / 1e12, peaks["bf16_tflops"], "TFLOP/s"
            kernels[name] = {"ms": round(avg, 4), "share": round(avg / ms_local, 4), "bound": bound,
                             "achieved": round(ach, 2), "peak": peak, "unit": unit, "frac": round(ach / peak, 4)}
            if name in ("fp1.three_nn", "fp2.three_nn") and getattr(eng, "overlap_geometry", False):
                # these searches run on a side stream beside the sampling of the next level: the main-stream event pair
                # only sees the wait for them, so no rate is derived from it
                kernels[name].update({"achieved": None, "frac": None, "overlapped": "side stream, beside sa%d.fps" %
                                      (3 - int(name[2]))})
        traffic = stage_traffic()
        for name, t in traffic.items():
            if name in kernels:
                kernels[name]["dram_bytes"] = t
        top = max(kernels, key=lambda k: kernels[k]["ms"])
        roof = {"kernel": top, "bound": kernels[top]["bound"], "achieved": kernels[top]["achieved"],
                "peak": kernels[top]["peak"], "unit": kernels[top]["unit"], "frac": kernels[top]["frac"],
                "traffic": traffic.get(top), "traffic_source": TRAFFIC_FILE if traffic.get(top) is not None else None,
                "peak_source": peaks["source"] + (" — sustained bf16 (kernel timed inside the step)"
                                                                   if kernels[top]["bound"] == "tensor" else "")}
        total_flops = sum(w for b, w in model.values() if b == "tensor")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.mlp_backend == "tcgen05" else "f32", "data": "synthetic",
            "config": {"workload": "PN2_CLS inference, synthetic tabletop clouds, 25600 points/scene (BASELINE config[1])",
                       "scenes_per_gpu_per_step": B, "parallelism": "scene-batch sharding, dp%d, no collective" % world,
                       "geometry": "fp32 exact (FPS / ball query / 3-NN)",
                       "mlp": "tcgen05 bf16 x bf16 -> fp32 (TMEM), BN folded" if args.mlp_backend == "tcgen05"
                       else "torch fp32 reference MLPs (NOT the product path)",
                       "l2": "flushed between steps (256 MiB memset outside the per-step event pairs)"},
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e},
            "gpu_launches": int(launches) * args.steps,
            "gpu_launches_per_step": int(launches),
            "roofline": roof,
            "model_tflops_per_step": round(total_flops / 1e12, 3),
            "model_tensor_frac": round(total_flops / (ms_local * 1e-3) / 1e12 / peaks["bf16_tflops"], 4),
            "kernels": kernels,
        }
        if world == 1 and not args.no_cpu_baseline:
            from oracle import model_cpu, pn2_ext_cpu
            torch.set_num_threads(os.cpu_count())
            sd = {k: v.cpu() for k, v in net.state_dict().items()}
            one = host_scenes[:1].contiguous()
            with torch.no_grad():
                model_cpu.pointnet2_forward(one, sd, model_cpu.PN2_CLS_CONFIG)  # warm-up
                ts = []
                for _ in range(2):
                    t0 = time.perf_counter()
                    model_cpu.pointnet2_forward(one, sd, model_cpu.PN2_CLS_CONFIG)
                    ts.append(time.perf_counter() - t0)
            line["cpu_baseline"] = {"value": 1.0 / min(ts), "unit": UNIT,
                                    "cores": max(pn2_ext_cpu.num_threads(), torch.get_num_threads()), "kind": "port",
                                    "sample": "1 scene of the batch, best of 2 after 1 warm-up (oracle/model_cpu.py)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
