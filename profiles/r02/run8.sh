#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -s 2>&1 | tail -40 > gpurun_out/r02/tests_train8.txt
grep -E "stage|passed|failed|FAILED|Error" gpurun_out/r02/tests_train8.txt | cut -c1-1800
