#!/bin/bash
# FPS pruned-cluster kernel: parity tests, per-phase profile, pipelined A/B against the full-scan kernel.
mkdir -p gpurun_out/fps3
O=gpurun_out/fps3
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fps" > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
tail -5 $O/tests.log
timeout 300 python tools/fps_profile.py > $O/fps_profile.txt 2>&1
cat $O/fps_profile.txt
timeout 600 python bench.py --steps 24 --warmup 6 --no-cpu-baseline --no-extras 2> $O/bench_pruned.err | grep '^{' | tail -1 > $O/bench_pruned.json
timeout 600 python bench.py --steps 24 --warmup 6 --no-cpu-baseline --no-extras --fps-full-scan 2> $O/bench_scan.err | grep '^{' | tail -1 > $O/bench_scan.json
python - <<'PY'
import json
for n in ("pruned", "scan"):
    try:
        d = json.load(open("gpurun_out/fps3/bench_%s.json" % n))
        print(n, d["ms_per_step"], d["value"], d.get("latency_ms_depth1"), d["kernels"]["layer1:fps"]["ms"], d["e2e"]["value"])
    except Exception as e:
        print(n, "failed", e)
PY
