#!/bin/bash
# round 2, call 42: fused training step and the module path (torch convolutions under autograd) on the SAME box
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 100 python bench_train.py --steps 5 --warmup 3 > $O/bt42_fused.json 2> $O/bt42.err
timeout 100 python bench_train.py --steps 3 --warmup 3 --path module > $O/bt42_module.json 2>> $O/bt42.err
python - <<PY
import json
for f in ("fused","module"):
    try:
        d=json.loads([l for l in open("$O/bt42_%s.json"%f) if l.startswith("{")][-1]); print(f, d["ms_per_step"], d["value"], d["peak_memory_GB"], d["clocks"]["sm_mhz"])
    except Exception as e: print(f, "failed", e)
PY
