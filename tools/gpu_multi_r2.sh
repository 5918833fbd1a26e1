#!/bin/bash
# usage: gpurun --gpus N --timeout 1200 -- bash tools/gpu_multi_r2.sh <tag> <N>
tag=$1; N=$2
out=gpurun_out/$tag; mkdir -p $out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== two-device test" >> $out/log.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k two_devices 2>&1 | tail -3 >> $out/log.txt
echo "== bench cfg2 x$N" >> $out/log.txt
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-extras > $out/bench_${N}gpu.json 2> $out/bench_${N}gpu.err; tail -2 $out/bench_${N}gpu.err >> $out/log.txt
python -c "import json; d=json.load(open('$out/bench_${N}gpu.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], 'e2e value', d['e2e']['value'])" >> $out/log.txt 2>&1
echo "== cfg3 x$N" >> $out/log.txt
timeout 300 $TR bench.py --workload cfg3 --gpus $N --steps 10 2>&1 | tail -1 > $out/cfg3_${N}gpu.json; cat $out/cfg3_${N}gpu.json >> $out/log.txt
echo "== cfg4 x$N" >> $out/log.txt
timeout 300 $TR bench.py --workload cfg4 --gpus $N --steps 4 2>&1 | tail -1 > $out/cfg4_${N}gpu.json; cat $out/cfg4_${N}gpu.json >> $out/log.txt
echo "== pcie probe x$N" >> $out/log.txt
timeout 200 $TR tools/pcie_probe.py 2>&1 | tail -$((N+2)) >> $out/log.txt
tail -c 6000 $out/log.txt
