#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python profiles/plan_crash_probe.py > gpurun_out/r02/plan_crash.txt 2>&1
timeout 600 python bench.py --no-reference-cuda > gpurun_out/r02/bench2.json 2> gpurun_out/r02/bench2.err
tail -3 gpurun_out/r02/plan_crash.txt
