#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_ops_gpu.py -m gpu -q 2>&1 | tail -6 > gpurun_out/r02/tests18.txt
timeout 600 python profiles/group_ab.py > gpurun_out/r02/group_ab.txt 2>&1
tail -3 gpurun_out/r02/tests18.txt; cat gpurun_out/r02/group_ab.txt
