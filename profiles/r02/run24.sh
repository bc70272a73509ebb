#!/bin/bash
# round 2, call 24: ncu --set full of the three GEMM epilogue flavours (plain / stats / bwd) at 2.6 M x 128 -> 128
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 600 ncu --set full --import-source on --clock-control none -k regex:gemm_bf16 -f -o $O/gemm_bwd python profiles/gemm_bwd_probe.py > $O/ncu24.log 2>&1
tail -3 $O/ncu24.log; ls -la $O/gemm_bwd.ncu-rep
