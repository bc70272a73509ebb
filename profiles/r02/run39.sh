#!/bin/bash
# round 2, call 39: ncu --set full of the training GEMM with 128 x 256 tiles at two K = 512 shapes (what bounds them)
mkdir -p gpurun_out/r02
O=gpurun_out/r02
S4G_PROBE_SHAPES="524288,512,1024;819200,512,256" timeout 600 ncu --set full --clock-control none -k regex:gemm_bf16 -f -o $O/gemm_tile256 python profiles/gemm_shape_probe.py > $O/ncu39.log 2>&1
tail -2 $O/ncu39.log; ls -la $O/gemm_tile256.ncu-rep
