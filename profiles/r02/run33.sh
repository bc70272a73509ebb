#!/bin/bash
# round 2, call 34: 128 x 256 tiles + two producer warps
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -x 2>&1 | tail -8 > $O/tests34.txt
tail -3 $O/tests34.txt
S4G_GEMM_LAYERS_TILES=1 timeout 600 python profiles/gemm_layers.py > $O/gemm_layers_v6.txt 2>&1; cat $O/gemm_layers_v6.txt
timeout 300 python bench_train.py --steps 8 --warmup 3 > $O/bt34.json 2> $O/bt34.err
timeout 300 python bench_train.py --steps 8 --warmup 3 > $O/bt34b.json 2>> $O/bt34.err
for f in bt34 bt34b; do python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/$f.json") if l.startswith("{")][-1]); print("$f", d["ms_per_step"], d["value"], d["peak_memory_GB"], d["loss"])
except Exception as e: print("$f", "failed", e)
PY
done
tail -3 $O/bt34.err
