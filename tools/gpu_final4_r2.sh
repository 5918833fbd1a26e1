#!/bin/bash
# Final round-2 evidence with the final build: parity suite, smoke, both bench arms, launch list of bench.py, FPS profile, probes
out=gpurun_out/final4; mkdir -p $out
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5 ) > $out/pytest.txt
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 ) > $out/smoke.txt
timeout 900 python bench.py --steps 32 --warmup 8 > $out/bench.json 2> $out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
timeout 300 python bench.py --steps 32 --warmup 8 --precision bf16 --no-extras --no-cpu-baseline 2>/dev/null | grep '^{' | tail -1 > $out/bench_bf16.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --depth 1 > /dev/null 2>&1
timeout 200 python tools/fps_profile.py > $out/fps_profile.txt 2>&1
( cd tools/probes && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/clc clc_probe.cu 2>/dev/null && timeout 60 /tmp/clc ) > $out/clc_probe.txt 2>&1
timeout 100 python tools/probes/d2h_split.py > $out/d2h_split.txt 2>&1
timeout 200 python tools/probes/order_probe.py 2>&1 | tail -2 > $out/order_probe.txt
timeout 400 python tools/ablate.py 8 > $out/ablate.txt 2>&1
cat $out/pytest.txt $out/smoke.txt; grep '^{' $out/bench.json | tail -1 | cut -c1-400; echo; grep '^{' $out/bench_reference.json | tail -1 | cut -c1-300; ls $out
