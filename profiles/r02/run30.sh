#!/bin/bash
# round 2, call 30: ncu --set full of the staged BatchNorm-backward kernels inside one training step (the last 24 launches: FP + SA levels)
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 900 ncu --profile-from-start off --set full --clock-control none -k 'regex:bn_bwd_reduce_staged|bn_bwd_apply_staged' --launch-skip 40 -c 24 -f -o $O/train_bwd2 python profiles/one_train_step.py > $O/ncu30.log 2>&1
tail -2 $O/ncu30.log; ls -la $O/train_bwd2.ncu-rep
