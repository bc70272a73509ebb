#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -s 2>&1 | tail -40 > gpurun_out/r02/tests_train9.txt
timeout 600 python bench_train.py --steps 3 --warmup 3 > gpurun_out/r02/bench_train9.json 2> gpurun_out/r02/bench_train9.err
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02/train9.csv python profiles/one_train_step.py > gpurun_out/r02/train9.log 2>&1
python profiles/one_train_step.py --summarize gpurun_out/r02/train9.csv > gpurun_out/r02/train9_summary.txt
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:gemm_bf16_kernel -c 40 -o gpurun_out/r02/train_gemm python profiles/one_train_step.py > gpurun_out/r02/train_gemm.log 2>&1
grep -E "stage|passed|failed|FAILED|Error" gpurun_out/r02/tests_train9.txt | cut -c1-1200; cut -c1-200 gpurun_out/r02/bench_train9.json; head -16 gpurun_out/r02/train9_summary.txt
