#!/bin/bash
# round 2, call 29: full GPU suite + smoke + benches of the state with the staged BatchNorm-backward passes; sanitizer on the new kernels
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/tests29.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke29.txt 2>&1
timeout 600 python bench.py > $O/bench29.json 2> $O/bench29.err
timeout 300 python bench_train.py --steps 8 --warmup 3 > $O/bt29.json 2> $O/bt29.err
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -x -k "bn_bwd_reduce_dense or bn_bwd_apply_against or pooled_reduce or sum_of_bf16 or head_logits or gemm_bwd_epilogue" 2>&1 | tail -25 > $O/sanitizer_memcheck_train2.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -x -k "bn_bwd_reduce_dense or bn_bwd_apply_against or gemm_epilogue_groups" 2>&1 | tail -25 > $O/sanitizer_racecheck_train2.txt
tail -3 $O/tests29.txt; tail -2 $O/smoke29.txt; cut -c1-250 $O/bench29.json; cut -c1-250 $O/bt29.json
for f in sanitizer_memcheck_train2 sanitizer_racecheck_train2; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard" $O/$f.txt | tail -5; done
