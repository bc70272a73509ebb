#!/bin/bash
# round 2, call 26: BatchNorm-backward apply staged by the copy engine; final-conv dW kernel with more blocks
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -x 2>&1 | tail -8 > $O/tests26.txt
tail -3 $O/tests26.txt
timeout 300 python bench_train.py --steps 5 --warmup 3 > $O/bt26_staged.json 2> $O/bt26.err
S4G_BWD_APPLY_VARIANT=1 timeout 300 python bench_train.py --steps 5 --warmup 3 > $O/bt26_regs.json 2>> $O/bt26.err
timeout 300 python bench_train.py --steps 5 --warmup 3 > $O/bt26_staged2.json 2>> $O/bt26.err
for f in staged regs staged2; do python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bt26_$f.json") if l.startswith("{")][-1]); print("$f", d["ms_per_step"], d["value"], d["peak_memory_GB"], d["loss"])
except Exception as e: print("$f", "failed", e)
PY
done
tail -3 $O/bt26.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/train26.csv python profiles/one_train_step.py > $O/ncu26.log 2>&1
python profiles/one_train_step.py --summarize $O/train26.csv > $O/train_kernels_v8.txt; head -24 $O/train_kernels_v8.txt
rm -f $O/train26.csv
