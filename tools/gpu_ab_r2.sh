#!/bin/bash
# generic A/B: bench.py with the flags given (one run per argument string), two rounds
mkdir -p gpurun_out/ab
O=gpurun_out/ab
i=0
for rep in 1 2; do
for f in "$@"; do
  i=$((i+1))
  timeout 600 python bench.py --steps 24 --warmup 6 --no-cpu-baseline --no-extras $f 2> $O/bench_$i.err | grep '^{' | tail -1 > $O/bench_$i.json
  python - "$i" "$f" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/ab/bench_%s.json" % sys.argv[1]))
    print("[%s] ms/step %.4f depth1 %.3f fps %.4f e2e %.1fM" % (sys.argv[2], d["ms_per_step"], d.get("latency_ms_depth1"), d["kernels"]["layer1:fps"]["ms"], d["e2e"]["value"] / 1e6))
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
done
done
