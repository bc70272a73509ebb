#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_train_engine_gpu.py -m gpu -q 2>&1 | tail -12 > gpurun_out/r02/tests_train16.txt
timeout 600 python bench_train.py --steps 5 --warmup 3 > gpurun_out/r02/bench_train16.json 2> gpurun_out/r02/bench_train16.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02/train16.csv python profiles/one_train_step.py > gpurun_out/r02/train16.log 2>&1
python profiles/one_train_step.py --summarize gpurun_out/r02/train16.csv > gpurun_out/r02/train16_summary.txt
tail -4 gpurun_out/r02/tests_train16.txt; cut -c1-200 gpurun_out/r02/bench_train16.json; head -8 gpurun_out/r02/train16_summary.txt; tail -3 gpurun_out/r02/bench_train16.err
