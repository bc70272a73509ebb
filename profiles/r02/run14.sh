#!/bin/bash
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q 2>&1 | tail -5 > gpurun_out/r02/tests14.txt
timeout 600 python bench.py --no-reference-cuda > gpurun_out/r02/bench14.json 2> gpurun_out/r02/bench14.err
timeout 300 python bench_large.py > gpurun_out/r02/bench_large14_c64.json 2> gpurun_out/r02/bench_large14_c64.err
timeout 300 python bench_large.py --chunk 74 --pool 74 > gpurun_out/r02/bench_large14_c74.json 2> gpurun_out/r02/bench_large14_c74.err
tail -3 gpurun_out/r02/tests14.txt; python - <<'PY'
import json
for f in ("bench14","bench_large14_c64","bench_large14_c74"):
    try:
        b=json.loads([l for l in open("gpurun_out/r02/%s.json"%f) if l.startswith("{")][-1])
        print(f, round(b["value"],1), round(b["ms_per_step"],3), b["e2e"]["value"], (b.get("e2e_bf16_host_buffers") or {}).get("value"))
    except Exception as e: print(f, "ERR", e, open("gpurun_out/r02/%s.err"%f).read()[-300:])
PY
