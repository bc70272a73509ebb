#!/bin/bash
mkdir -p gpurun_out/r02
cd /root/repo
timeout 1200 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -x -k "gemm_bf16 or colstats or head_logits or group_rows or interp_rows or block_forward or dropout" 2>&1 | tail -25 > gpurun_out/r02/sanitizer_memcheck_train.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -x -k "test_gemm_bf16_fused_statistics or test_gemm_weight_stationary" 2>&1 | tail -25 > gpurun_out/r02/sanitizer_racecheck_gemm.txt
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_chain_gpu.py -m gpu -q -x -k "linear_tf32 or one_block_chain" 2>&1 | tail -15 > gpurun_out/r02/sanitizer_memcheck_tf32.txt
timeout 600 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "staged or gather_points_kernel" 2>&1 | tail -15 > gpurun_out/r02/sanitizer_memcheck_group.txt
for f in sanitizer_memcheck_train sanitizer_racecheck_gemm sanitizer_memcheck_tf32 sanitizer_memcheck_group; do echo "== $f"; grep -E "ERROR SUMMARY|passed|failed|Error|hazard" gpurun_out/r02/$f.txt | tail -5; done
