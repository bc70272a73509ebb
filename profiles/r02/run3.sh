#!/bin/bash
mkdir -p gpurun_out/r02
for a in "0 0" "0 1" "1 1"; do
  echo "=== tiny_repro $a" >> gpurun_out/r02/tiny.txt
  CUDA_LAUNCH_BLOCKING=1 timeout 300 python profiles/tiny_repro.py $a >> gpurun_out/r02/tiny.txt 2>&1
  echo "rc=$?" >> gpurun_out/r02/tiny.txt
done
echo "=== sanitizer 0 1" >> gpurun_out/r02/tiny.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python profiles/tiny_repro.py 0 1 2>&1 | grep -v "^$" | tail -60 >> gpurun_out/r02/tiny.txt
S4G_PARITY_REPORT=gpurun_out/r02/parity2.json timeout 900 python -m pytest tests/test_pose_parity_gpu.py -m gpu -q -s 2>&1 | tail -40 > gpurun_out/r02/tests_parity2.txt
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_pose_parity_gpu.py 2>&1 | tail -60 > gpurun_out/r02/tests3.txt
