#!/bin/bash
# round 2, call 32: second TMA producer warp in the training GEMM
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -x 2>&1 | tail -8 > $O/tests32.txt
tail -3 $O/tests32.txt
timeout 600 python profiles/gemm_layers.py > $O/gemm_layers_v4.txt 2>&1; cat $O/gemm_layers_v4.txt
timeout 300 python bench_train.py --steps 8 --warmup 3 > $O/bt32.json 2> $O/bt32.err
timeout 300 python bench_train.py --steps 8 --warmup 3 > $O/bt32b.json 2>> $O/bt32.err
for f in bt32 bt32b; do python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/$f.json") if l.startswith("{")][-1]); print("$f", d["ms_per_step"], d["value"], d["peak_memory_GB"], d["loss"])
except Exception as e: print("$f", "failed", e)
PY
done
tail -3 $O/bt32.err
