#!/bin/bash
# round 2, call 31: pooled BatchNorm-backward apply without the 64-bit division / per-channel selects; max-pool with loads in flight
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -x 2>&1 | tail -8 > $O/tests31.txt
tail -3 $O/tests31.txt
timeout 300 python bench_train.py --steps 8 --warmup 3 > $O/bt31.json 2> $O/bt31.err
timeout 300 python bench_train.py --steps 8 --warmup 3 > $O/bt31b.json 2>> $O/bt31.err
for f in bt31 bt31b; do python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/$f.json") if l.startswith("{")][-1]); print("$f", d["ms_per_step"], d["value"], d["peak_memory_GB"], d["loss"])
except Exception as e: print("$f", "failed", e)
PY
done
tail -3 $O/bt31.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/train31.csv python profiles/one_train_step.py > $O/ncu31.log 2>&1
python profiles/one_train_step.py --summarize $O/train31.csv > $O/train_kernels_v9.txt; head -12 $O/train_kernels_v9.txt
grep "apply_staged\|maxpool" $O/train31.csv | awk -F'","' '{print $5, $(NF)}' | sed 's/(.*)//' | sort -k2 -n -r -t' ' | head -8
rm -f $O/train31.csv
