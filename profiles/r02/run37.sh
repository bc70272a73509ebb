#!/bin/bash
# round 2, call 37: full GPU suite + smoke + benches of the final state; sanitizer on the kernels added since call 29
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > $O/tests37.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke37.txt 2>&1
timeout 600 python bench.py > $O/bench37.json 2> $O/bench37.err
timeout 300 python bench_train.py --steps 8 --warmup 3 > $O/bt37.json 2> $O/bt37.err
timeout 900 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -x -k "interp_rows_backward or group_rows_and_scatter or head_logits or gemm_256_column or block_forward_backward or bn_bwd_apply_against" 2>&1 | tail -25 > $O/sanitizer_memcheck_train3.txt
timeout 900 compute-sanitizer --tool racecheck --print-limit 10 python -m pytest tests/test_train_engine_gpu.py -m gpu -q -x -k "gemm_256_column or gemm_epilogue_groups or head_logits" 2>&1 | tail -25 > $O/sanitizer_racecheck_train3.txt
tail -3 $O/tests37.txt; tail -2 $O/smoke37.txt; cut -c1-250 $O/bench37.json; cut -c1-250 $O/bt37.json
for f in sanitizer_memcheck_train3 sanitizer_racecheck_train3; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|Error|hazard" $O/$f.txt | tail -5; done
