#!/bin/bash
mkdir -p gpurun_out/r02
rm -f gpurun_out/r02/parity3.json
S4G_PARITY_REPORT=gpurun_out/r02/parity3.json timeout 900 python -m pytest tests/test_pose_parity_gpu.py -m gpu -q -s 2>&1 | tail -40 > gpurun_out/r02/tests_parity3.txt
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_pose_parity_gpu.py 2>&1 | tail -60 > gpurun_out/r02/tests4.txt
timeout 600 python bench.py > gpurun_out/r02/bench4.json 2> gpurun_out/r02/bench4.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,launch__registers_per_thread,launch__grid_size,launch__block_size
timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02/launches4.csv python profiles/one_forward.py > gpurun_out/r02/launches4.log 2>&1
tail -3 gpurun_out/r02/tests4.txt
