#!/bin/bash
# round 2, call 27: finest propagation level with the conv on the sparse rows (training)
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python -m pytest tests/test_train_engine_gpu.py tests/test_train_gpu.py -m gpu -q -x 2>&1 | tail -8 > $O/tests27.txt
tail -3 $O/tests27.txt
timeout 300 python bench_train.py --steps 5 --warmup 3 > $O/bt27_split.json 2> $O/bt27.err
timeout 300 python bench_train.py --steps 5 --warmup 3 --no-fp-split > $O/bt27_nosplit.json 2>> $O/bt27.err
timeout 300 python bench_train.py --steps 5 --warmup 3 > $O/bt27_split2.json 2>> $O/bt27.err
for f in split nosplit split2; do python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bt27_$f.json") if l.startswith("{")][-1]); print("$f", d["ms_per_step"], d["value"], d["peak_memory_GB"], d["loss"], d["e2e"]["value"])
except Exception as e: print("$f", "failed", e)
PY
done
tail -3 $O/bt27.err
