#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_train_engine_gpu.py -m gpu -q 2>&1 | tail -6 > gpurun_out/r02/tests_train17.txt
timeout 600 python bench_train.py --steps 5 --warmup 3 > gpurun_out/r02/bench_train17_a.json 2> gpurun_out/r02/bench_train17_a.err
S4G_BWD_REDUCE_VARIANT=1 timeout 600 python bench_train.py --steps 5 --warmup 3 > gpurun_out/r02/bench_train17_b.json 2> gpurun_out/r02/bench_train17_b.err
timeout 900 ncu --profile-from-start off --set full --clock-control none -k 'regex:bn_bwd' -s 20 -c 8 -o gpurun_out/r02/train_bwd_ncu python profiles/one_train_step.py > gpurun_out/r02/train_bwd_ncu.log 2>&1
tail -3 gpurun_out/r02/tests_train17.txt; cut -c1-190 gpurun_out/r02/bench_train17_a.json; cut -c1-190 gpurun_out/r02/bench_train17_b.json
