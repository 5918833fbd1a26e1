#!/bin/bash
# usage: gpurun --gpus 8 --timeout 1200 -- bash tools/gpu_scale2_r2.sh <tag> <N>   final build on N GPUs: bench, config 4 with peer-memory and NCCL moments (same box)
tag=$1; N=$2
out=gpurun_out/$tag; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-extras 2> $out/bench_${N}gpu.err | grep '^{' | tail -1 > $out/bench_${N}gpu.json
python -c "import json; d=json.load(open('$out/bench_${N}gpu.json')); print('N=$N value', d['value'], 'ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'e2e value', d['e2e']['value'])"
timeout 300 $TR bench.py --workload cfg4 --gpus $N --steps 6 2>$out/cfg4_peer.err | grep '^{' | tail -1 | tee $out/cfg4_${N}gpu_peer_moments.json | cut -c1-330
timeout 300 $TR bench.py --workload cfg4 --gpus $N --steps 6 --nccl-moments 2>$out/cfg4_nccl.err | grep '^{' | tail -1 | tee $out/cfg4_${N}gpu_nccl_moments.json | cut -c1-330
