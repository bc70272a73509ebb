#!/bin/bash
# round 2, call 23: BWD epilogue with two y tiles, pipelined column walk
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -x 2>&1 | tail -8 > $O/tests23.txt
tail -3 $O/tests23.txt
timeout 300 python bench_train.py --steps 5 --warmup 3 --no-fused-bwd-reduce > $O/bt23_nofuse.json 2> $O/bt23.err
timeout 300 python bench_train.py --steps 5 --warmup 3 > $O/bt23_all.json 2>> $O/bt23.err
for f in all nofuse; do python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bt23_$f.json") if l.startswith("{")][-1]); print("$f", d["ms_per_step"], d["value"], d["peak_memory_GB"], d["loss"])
except Exception as e: print("$f", "failed", e)
PY
done
tail -3 $O/bt23.err
timeout 600 python profiles/gemm_layers.py 2>&1 | grep "dX\|rows" > $O/gemm_layers_v3.txt; cat $O/gemm_layers_v3.txt

