#!/bin/bash
# round 2: the training step at 1 and 8 GPUs of one box (weak scaling, NCCL gradient all-reduce) — final state of the round
mkdir -p gpurun_out/r02
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python bench_train.py --steps 8 --warmup 3 > gpurun_out/r02/m4_bench_train_1gpu.json 2> gpurun_out/r02/m4_bench_train_1gpu.err
timeout 400 $TR --nproc-per-node 8 --master-port 29811 bench_train.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r02/m4_bench_train_8gpu.json 2> gpurun_out/r02/m4_bench_train_8gpu.err
python - <<'PY'
import json
for f in ("m4_bench_train_1gpu","m4_bench_train_8gpu"):
    try:
        b=json.loads([l for l in open("gpurun_out/r02/%s.json"%f) if l.startswith("{")][-1])
        print(f, round(b["value"],1), round(b["ms_per_step"],3), round(b["e2e"]["value"],1), b.get("peak_memory_GB"))
    except Exception as e: print(f, "ERR", e, open("gpurun_out/r02/%s.err"%f).read()[-300:])
PY
