#!/bin/bash
# One gpurun box visit: parity tests, chain phase profile, op timings, bench A/B over the tuning doors.
# usage: gpurun --timeout 1500 -- bash tools/gpu_session.sh <tag>
tag=${1:-s}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $out/pytest.txt
echo "== default" > $out/tc_profile.txt
timeout 200 python tools/tc_profile.py >> $out/tc_profile.txt 2>&1
echo "== GSPN_TC_OCC=1" >> $out/tc_profile.txt
GSPN_TC_OCC=1 timeout 200 python tools/tc_profile.py >> $out/tc_profile.txt 2>&1
timeout 300 python tools/op_bench.py mlp interp > $out/op_bench_default.json 2> $out/op_bench_default.err
GSPN_TC_OCC=1 timeout 300 python tools/op_bench.py mlp interp > $out/op_bench_occ1.json 2> $out/op_bench_occ1.err
timeout 400 python bench.py --steps 24 --warmup 6 > $out/bench_default.json 2> $out/bench_default.err
GSPN_TC_OCC=1 timeout 400 python bench.py --steps 24 --warmup 6 --no-cpu-baseline > $out/bench_occ1.json 2> $out/bench_occ1.err
GSPN_FPS_CFG=256,32,4 timeout 400 python bench.py --steps 24 --warmup 6 --no-cpu-baseline > $out/bench_fps4.json 2> $out/bench_fps4.err
tail -3 $out/pytest.txt; cat $out/tc_profile.txt; for f in $out/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["e2e"]["ms_per_step"], {k:v["ms"] for k,v in list(d["kernels"].items())[:12]})
except Exception as e:
    print("bad", e)
PY
done
