#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -s 2>&1 | tail -40 > gpurun_out/r02/tests_train6.txt
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02/train6.csv python profiles/one_train_step.py > gpurun_out/r02/train6.log 2>&1
python profiles/one_train_step.py --summarize gpurun_out/r02/train6.csv > gpurun_out/r02/train6_summary.txt
tail -8 gpurun_out/r02/tests_train6.txt; head -45 gpurun_out/r02/train6_summary.txt
