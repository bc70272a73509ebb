#!/bin/bash
# round 2, call 36: grouping gradient as a gather, head-logit kernels on row tiles
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -x 2>&1 | tail -8 > $O/tests36.txt
tail -3 $O/tests36.txt
timeout 300 python bench_train.py --steps 8 --warmup 3 > $O/bt36.json 2> $O/bt36.err
timeout 300 python bench_train.py --steps 8 --warmup 3 --interp-bwd-scatter > $O/bt36_scatter.json 2>> $O/bt36.err
timeout 300 python bench_train.py --steps 8 --warmup 3 > $O/bt36b.json 2>> $O/bt36.err
for f in bt36 bt36_scatter bt36b; do python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/$f.json") if l.startswith("{")][-1]); print("$f", d["ms_per_step"], d["value"], d["peak_memory_GB"], d["loss"])
except Exception as e: print("$f", "failed", e)
PY
done
tail -3 $O/bt36.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/train36.csv python profiles/one_train_step.py > $O/ncu36.log 2>&1
python profiles/one_train_step.py --summarize $O/train36.csv > $O/train_kernels_v11.txt; head -30 $O/train_kernels_v11.txt
rm -f $O/train36.csv
