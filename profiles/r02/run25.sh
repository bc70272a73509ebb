#!/bin/bash
# round 2, call 25: BatchNorm-backward reduce staged through shared memory by the copy engine
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -x 2>&1 | tail -8 > $O/tests25.txt
tail -3 $O/tests25.txt
timeout 300 python bench_train.py --steps 5 --warmup 3 > $O/bt25_staged.json 2> $O/bt25.err
S4G_BWD_REDUCE_VARIANT=2 timeout 300 python bench_train.py --steps 5 --warmup 3 > $O/bt25_regs.json 2>> $O/bt25.err
timeout 300 python bench_train.py --steps 5 --warmup 3 > $O/bt25_staged2.json 2>> $O/bt25.err
for f in staged regs staged2; do python - <<PY
import json
try:
    d=json.loads([l for l in open("$O/bt25_$f.json") if l.startswith("{")][-1]); print("$f", d["ms_per_step"], d["value"], d["peak_memory_GB"], d["loss"])
except Exception as e: print("$f", "failed", e)
PY
done
tail -3 $O/bt25.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/train25.csv python profiles/one_train_step.py > $O/ncu25.log 2>&1
python profiles/one_train_step.py --summarize $O/train25.csv > $O/train_kernels_v7.txt; head -24 $O/train_kernels_v7.txt
grep "bn_bwd_reduce" $O/train25.csv | awk -F'","' '{print $5, $(NF)}' | head -40 > $O/reduce_launches.txt
rm -f $O/train25.csv
