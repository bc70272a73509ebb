#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -s 2>&1 | tail -40 > gpurun_out/r02/tests_train7.txt
timeout 600 python bench_train.py --steps 3 --warmup 3 > gpurun_out/r02/bench_train7.json 2> gpurun_out/r02/bench_train7.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02/train7.csv python profiles/one_train_step.py > gpurun_out/r02/train7.log 2>&1
python profiles/one_train_step.py --summarize gpurun_out/r02/train7.csv > gpurun_out/r02/train7_summary.txt
grep -E "stage|passed|failed|FAILED|Error" gpurun_out/r02/tests_train7.txt | cut -c1-1500; cut -c1-200 gpurun_out/r02/bench_train7.json; head -24 gpurun_out/r02/train7_summary.txt
