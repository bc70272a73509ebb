#!/bin/bash
# round 2, GPU call 1: full GPU suite (+ parity report), bench lines, ncu launch list, ncu --set full of the geometry kernels
mkdir -p gpurun_out/r02
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02/smi.txt
S4G_PARITY_REPORT=gpurun_out/r02/parity.json timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_pose_parity_gpu.py 2>&1 | tail -25 > gpurun_out/r02/tests1.txt
S4G_PARITY_REPORT=gpurun_out/r02/parity.json timeout 600 python -m pytest tests/test_pose_parity_gpu.py -m gpu -q -s 2>&1 | tail -40 > gpurun_out/r02/tests_parity.txt
timeout 300 python profiles/dump_tuned_plans.py > gpurun_out/r02/tuned_plans_run.json 2> gpurun_out/r02/tuned_plans_run.err
timeout 600 python bench.py > gpurun_out/r02/bench1.json 2> gpurun_out/r02/bench1.err
timeout 300 python bench.py --no-fp-split --no-cpu-baseline --no-reference-cuda > gpurun_out/r02/bench1_nosplit.json 2> gpurun_out/r02/bench1_nosplit.err
timeout 300 python bench.py --batch 1 --no-cpu-baseline --no-reference-cuda > gpurun_out/r02/bench1_b1.json 2> gpurun_out/r02/bench1_b1.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,launch__registers_per_thread,launch__grid_size,launch__block_size
timeout 600 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r02/launches.csv python profiles/one_forward.py > gpurun_out/r02/launches.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:fps_kernel|ball_query|three_nn_grid|interp_concat' -o gpurun_out/r02/geom python profiles/one_forward.py > gpurun_out/r02/geom.log 2>&1
ls -la gpurun_out/r02
