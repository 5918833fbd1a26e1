#!/bin/bash
# dynamic tile scheduling of the chain kernels: parity suite, then pipelined A/B against the static stride
mkdir -p gpurun_out/dyn
O=gpurun_out/dyn
timeout 1500 python -m pytest tests -x -q -m gpu > $O/tests.log 2>&1; echo "tests rc=$?" >> $O/tests.log
tail -4 $O/tests.log
for v in dyn static; do
  f=""; [ $v = static ] && f="--static-tiles"
  timeout 600 python bench.py --steps 24 --warmup 6 --no-cpu-baseline --no-extras $f 2> $O/bench_$v.err | grep '^{' | tail -1 > $O/bench_$v.json
done
timeout 300 python tools/chain_only.py > $O/chain_only.txt 2>&1
python - <<'PY'
import json
for n in ("dyn", "static"):
    try:
        d = json.load(open("gpurun_out/dyn/bench_%s.json" % n))
        k = d["kernels"]
        print(n, "ms/step %.4f" % d["ms_per_step"], "depth1 %.3f" % d.get("latency_ms_depth1"), "e2e %.3f" % d["e2e"]["ms_per_step"] if "ms_per_step" in d["e2e"] else d["e2e"]["value"],
              " ".join("%s=%.4f" % (s, k[s]["ms"]) for s in ("layer1:mlp", "layer2:mlp", "layer3:mlp", "layer4:mlp", "fa_layer4:mlp", "fa_layer1:mlp")))
    except Exception as e:
        print(n, "failed", e)
PY
tail -12 $O/chain_only.txt
