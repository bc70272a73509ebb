#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_train_engine_gpu.py -m gpu -q 2>&1 | tail -14 > gpurun_out/r02/tests_train12.txt
timeout 600 python bench_train.py --steps 5 --warmup 3 > gpurun_out/r02/bench_train12.json 2> gpurun_out/r02/bench_train12.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02/train12.csv python profiles/one_train_step.py > gpurun_out/r02/train12.log 2>&1
python profiles/one_train_step.py --summarize gpurun_out/r02/train12.csv > gpurun_out/r02/train12_summary.txt
tail -6 gpurun_out/r02/tests_train12.txt; cut -c1-200 gpurun_out/r02/bench_train12.json; grep -o '"peak_memory_GB": [0-9.]*' gpurun_out/r02/bench_train12.json; head -14 gpurun_out/r02/train12_summary.txt; tail -3 gpurun_out/r02/bench_train12.err
