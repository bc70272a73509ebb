#!/bin/bash
# round 2, call 20: GEMM with two epilogue groups, BatchNorm-backward reduce fused into the input-gradient GEMM, pooled reduce
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 900 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -x 2>&1 | tail -15 > $O/tests20.txt
tail -5 $O/tests20.txt
timeout 300 python bench_train.py --steps 5 --warmup 3 > $O/bt20_all.json 2> $O/bt20_all.err
timeout 300 python bench_train.py --steps 5 --warmup 3 --epilogue-groups 1 > $O/bt20_g1.json 2>> $O/bt20_all.err
timeout 300 python bench_train.py --steps 5 --warmup 3 --no-fused-bwd-reduce > $O/bt20_nofuse.json 2>> $O/bt20_all.err
timeout 300 python bench_train.py --steps 5 --warmup 3 --no-sparse-pool-reduce > $O/bt20_nosparse.json 2>> $O/bt20_all.err
for f in all g1 nofuse nosparse; do python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bt20_$f.json") if l.startswith("{")][-1]); print("$f", d["ms_per_step"], d["value"], d["peak_memory_GB"], d["loss"])
except Exception as e: print("$f", "failed", e)
PY
done
tail -5 $O/bt20_all.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/train20.csv python profiles/one_train_step.py > $O/ncu20.log 2>&1
python profiles/one_train_step.py --summarize $O/train20.csv > $O/train_kernels_v6.txt; head -30 $O/train_kernels_v6.txt
rm -f $O/train20.csv
timeout 300 python profiles/train_torch_ops.py > $O/train_torch_ops.txt 2>&1; head -50 $O/train_torch_ops.txt | cut -c1-220
