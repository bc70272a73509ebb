#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -x -s 2>&1 | tail -60 > gpurun_out/r02/tests_train5.txt
timeout 600 python bench_train.py --steps 3 --warmup 3 > gpurun_out/r02/bench_train5.json 2> gpurun_out/r02/bench_train5.err
tail -5 gpurun_out/r02/tests_train5.txt; tail -3 gpurun_out/r02/bench_train5.err; cat gpurun_out/r02/bench_train5.json | cut -c1-600
