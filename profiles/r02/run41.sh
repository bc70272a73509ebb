#!/bin/bash
# round 2, call 41: training tests + bench after the GEMM launch planning moved into plan_gemm()
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 400 python -m pytest tests/test_train_engine_gpu.py tests/test_train_gpu.py -m gpu -q -x 2>&1 | tail -3 > $O/tests41.txt
tail -1 $O/tests41.txt
timeout 200 python bench_train.py --steps 6 --warmup 3 > $O/bt41.json 2> $O/bt41.err
python - <<PY
import json
d=json.loads([l for l in open("$O/bt41.json") if l.startswith("{")][-1]); print(d["ms_per_step"], d["value"], d["clocks"]["sm_mhz"])
PY
