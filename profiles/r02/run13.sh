#!/bin/bash
mkdir -p gpurun_out/r02
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02/tests13.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02/smoke13.txt 2>&1
timeout 600 python bench.py > gpurun_out/r02/bench13.json 2> gpurun_out/r02/bench13.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02/bench13_ref.json 2> gpurun_out/r02/bench13_ref.err
tail -5 gpurun_out/r02/tests13.txt; cat gpurun_out/r02/smoke13.txt | tail -2; cut -c1-300 gpurun_out/r02/bench13.json; cut -c1-400 gpurun_out/r02/bench13_ref.json
