#!/bin/bash
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r02/tests15.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02/smoke15.txt 2>&1
timeout 600 python bench.py > gpurun_out/r02/bench15.json 2> gpurun_out/r02/bench15.err
timeout 900 ncu --profile-from-start off --set full --clock-control none -k 'regex:gemm_bf16|bn_bwd|bn_act' -c 24 -o gpurun_out/r02/train_ncu python profiles/one_train_step.py > gpurun_out/r02/train_ncu.log 2>&1
ls -la gpurun_out/r02/train_ncu.ncu-rep
tail -3 gpurun_out/r02/tests15.txt; tail -1 gpurun_out/r02/smoke15.txt; cut -c1-160 gpurun_out/r02/bench15.json
