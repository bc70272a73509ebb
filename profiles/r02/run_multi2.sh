#!/bin/bash
mkdir -p gpurun_out/r02
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 python bench_train.py --steps 5 --warmup 3 > gpurun_out/r02/m2_bench_train_1gpu.json 2> gpurun_out/r02/m2_bench_train_1gpu.err
timeout 400 $TR --nproc-per-node 8 --master-port 29611 bench_train.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02/m2_bench_train_8gpu.json 2> gpurun_out/r02/m2_bench_train_8gpu.err
timeout 300 python bench_large.py --gpus 1 > gpurun_out/r02/m2_bench_large_1gpu.json 2> gpurun_out/r02/m2_bench_large_1gpu.err
for n in 2 4 8; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29620+n)) bench_large.py --gpus $n > gpurun_out/r02/m2_bench_large_${n}gpu.json 2> gpurun_out/r02/m2_bench_large_${n}gpu.err
done
timeout 300 python bench.py --no-cpu-baseline --no-reference-cuda > gpurun_out/r02/m2_bench_1gpu.json 2> gpurun_out/r02/m2_bench_1gpu.err
timeout 300 $TR --nproc-per-node 8 --master-port 29640 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02/m2_bench_8gpu.json 2> gpurun_out/r02/m2_bench_8gpu.err
python - <<'PY'
import json
for f in ("m2_bench_train_1gpu","m2_bench_train_8gpu","m2_bench_large_1gpu","m2_bench_large_2gpu","m2_bench_large_4gpu","m2_bench_large_8gpu","m2_bench_1gpu","m2_bench_8gpu"):
    try:
        b=json.loads([l for l in open("gpurun_out/r02/%s.json"%f) if l.startswith("{")][-1])
        print(f, round(b["value"],1), round(b["ms_per_step"],3), round(b["e2e"]["value"],1), (b.get("e2e_bf16_host_buffers") or {}).get("value"), b.get("peak_memory_GB"))
    except Exception as e: print(f, "ERR", e, open("gpurun_out/r02/%s.err"%f).read()[-300:])
PY
