#!/bin/bash
# round 2, call 22: BWD epilogue v2 (y tile + column walk), per-launch epilogue groups
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -x 2>&1 | tail -8 > $O/tests22.txt
tail -3 $O/tests22.txt
timeout 300 python bench_train.py --steps 5 --warmup 3 --no-fused-bwd-reduce > $O/bt22_nofuse.json 2> $O/bt22.err
timeout 300 python bench_train.py --steps 5 --warmup 3 > $O/bt22_all.json 2>> $O/bt22.err
for f in all nofuse; do python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bt22_$f.json") if l.startswith("{")][-1]); print("$f", d["ms_per_step"], d["value"], d["peak_memory_GB"], d["loss"])
except Exception as e: print("$f", "failed", e)
PY
done
tail -3 $O/bt22.err
timeout 600 python profiles/gemm_layers.py > $O/gemm_layers.txt 2>&1; cat $O/gemm_layers.txt
