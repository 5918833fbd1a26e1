#!/bin/bash
# FPS with two clouds per CTA: parity, then pipelined A/B
mkdir -p gpurun_out/pack
O=gpurun_out/pack
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fps or dynamic_tiles" > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
tail -4 $O/tests.log
for v in 1 2 1 2; do
  timeout 600 python bench.py --steps 24 --warmup 6 --no-cpu-baseline --no-extras --fps-pack $v 2> $O/bench_$v.err | grep '^{' | tail -1 > $O/bench_$v.json
  python - $v <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.load(open("gpurun_out/pack/bench_%s.json" % v))
    print("pack", v, "ms/step %.4f" % d["ms_per_step"], "depth1 %.3f" % d.get("latency_ms_depth1"), "fps %.4f" % d["kernels"]["layer1:fps"]["ms"], "e2e", d["e2e"]["value"])
except Exception as e:
    print(v, "failed", e)
PY
done
