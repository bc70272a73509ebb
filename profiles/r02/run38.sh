#!/bin/bash
# round 2, call 38: staged BatchNorm-backward apply specialised at compile time (dense / pooled, dropout / none)
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -x 2>&1 | tail -5 > $O/tests38.txt
tail -2 $O/tests38.txt
timeout 300 python bench_train.py --steps 8 --warmup 3 > $O/bt38.json 2> $O/bt38.err
python - <<PY
import json
d=json.loads([l for l in open("$O/bt38.json") if l.startswith("{")][-1]); print(d["ms_per_step"], d["value"], d["clocks"])
PY
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/train38.csv python profiles/one_train_step.py > $O/ncu38.log 2>&1
python profiles/one_train_step.py --summarize $O/train38.csv > $O/train_kernels_v12.txt; head -8 $O/train_kernels_v12.txt
rm -f $O/train38.csv
