#!/bin/bash
# round 2, call 35: interpolation gradient as a gather over the inverted 3-NN index
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -x 2>&1 | tail -8 > $O/tests35.txt
tail -3 $O/tests35.txt
timeout 300 python bench_train.py --steps 8 --warmup 3 > $O/bt35.json 2> $O/bt35.err
timeout 300 python bench_train.py --steps 8 --warmup 3 --interp-bwd-scatter > $O/bt35_scatter.json 2>> $O/bt35.err
timeout 300 python bench_train.py --steps 8 --warmup 3 > $O/bt35b.json 2>> $O/bt35.err
for f in bt35 bt35_scatter bt35b; do python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/$f.json") if l.startswith("{")][-1]); print("$f", d["ms_per_step"], d["value"], d["peak_memory_GB"], d["loss"])
except Exception as e: print("$f", "failed", e)
PY
done
tail -3 $O/bt35.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/train35.csv python profiles/one_train_step.py > $O/ncu35.log 2>&1
python profiles/one_train_step.py --summarize $O/train35.csv > $O/train_kernels_v10.txt; head -30 $O/train_kernels_v10.txt
rm -f $O/train35.csv
