#!/bin/bash
# usage: gpurun --gpus 8 --timeout 1500 -- bash tools/gpu_scale_r2.sh <tag> <N>     bench + concurrent D2H probe on N GPUs of one box
tag=$1; N=$2
out=gpurun_out/$tag; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-extras 2> $out/bench_${N}gpu.err | grep '^{' | tail -1 > $out/bench_${N}gpu.json
python -c "import json; d=json.load(open('$out/bench_${N}gpu.json')); print('N=$N value', d['value'], 'ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'e2e value', d['e2e']['value'])"
timeout 200 $TR tools/pcie_probe.py 2>/dev/null | tail -$((N+2)) | tee $out/pcie_probe_${N}gpu.txt
timeout 300 $TR bench.py --workload cfg3 --gpus $N --steps 20 2>/dev/null | grep '^{' | tail -1 | tee $out/cfg3_${N}gpu.json | cut -c1-260
timeout 300 $TR bench.py --workload cfg4 --gpus $N --steps 6 2>/dev/null | grep '^{' | tail -1 | tee $out/cfg4_${N}gpu.json | cut -c1-330
