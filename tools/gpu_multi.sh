#!/bin/bash
# usage: gpurun --gpus N --timeout 900 -- bash tools/gpu_multi.sh N
n=$1; out=gpurun_out/multi; mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 24 --warmup 6 > $out/bench_${n}gpu.json 2> $out/bench_${n}gpu.err
tail -c 1500 $out/bench_${n}gpu.json | cut -c1-700; tail -3 $out/bench_${n}gpu.err
