#!/bin/bash
# round 2, call 28: ncu --set full of the training GEMM at a wide-N shape (0.6 of the copy peak) next to a narrow one (0.97)
mkdir -p gpurun_out/r02
O=gpurun_out/r02
S4G_PROBE_SHAPES="2097152,256,512;2621440,128,256;524288,512,1024" timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_bf16 -f -o $O/gemm_shapes python profiles/gemm_shape_probe.py > $O/ncu28.log 2>&1
tail -2 $O/ncu28.log; ls -la $O/gemm_shapes.ncu-rep
