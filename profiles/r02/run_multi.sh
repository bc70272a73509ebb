#!/bin/bash
# 8-GPU box: training step (NCCL gradient all-reduce) at 1 / 8 GPUs, large-batch stream at 1 / 2 / 4 / 8, bench.py e2e at 8
mkdir -p gpurun_out/r02
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/r02/topo.txt 2>&1
timeout 300 python bench_train.py --steps 5 --warmup 3 > gpurun_out/r02/bench_train_1gpu.json 2> gpurun_out/r02/bench_train_1gpu.err
timeout 400 $TR --nproc-per-node 8 --master-port 29511 bench_train.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02/bench_train_8gpu.json 2> gpurun_out/r02/bench_train_8gpu.err
timeout 400 $TR --nproc-per-node 8 --master-port 29512 bench_train.py --gpus 8 --steps 5 --warmup 3 --path module > gpurun_out/r02/bench_train_8gpu_module.json 2> gpurun_out/r02/bench_train_8gpu_module.err
timeout 300 python bench_large.py --gpus 1 > gpurun_out/r02/bench_large_1gpu.json 2> gpurun_out/r02/bench_large_1gpu.err
for n in 2 4 8; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29520+n)) bench_large.py --gpus $n > gpurun_out/r02/bench_large_${n}gpu.json 2> gpurun_out/r02/bench_large_${n}gpu.err
done
timeout 300 $TR --nproc-per-node 8 --master-port 29540 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02/bench_8gpu.json 2> gpurun_out/r02/bench_8gpu.err
for f in bench_train_1gpu bench_train_8gpu bench_train_8gpu_module bench_large_1gpu bench_large_2gpu bench_large_4gpu bench_large_8gpu bench_8gpu; do echo "== $f"; cut -c1-260 gpurun_out/r02/$f.json; tail -2 gpurun_out/r02/$f.err | cut -c1-300; done
