#!/bin/bash
# Round-end evidence on one B200: parity tests, smoke, bench (both arms), ncu launch list + full capture, ablation.
out=gpurun_out/final; mkdir -p $out
( timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 ) > $out/pytest.txt
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 ) > $out/smoke.txt
timeout 600 python bench.py --steps 32 --warmup 8 > $out/bench.json 2> $out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $out/launches.csv python tools/one_forward.py > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o $out/fwd python tools/one_forward.py > $out/ncu_full.log 2>&1
timeout 300 python tools/ablate.py 8 > $out/ablate.txt 2>&1
timeout 100 python tools/pcie_probe.py > $out/pcie.txt 2>&1
timeout 200 python tools/tc_profile.py > $out/tc_profile.txt 2>&1
cat $out/pytest.txt $out/smoke.txt; tail -c 600 $out/bench.json; echo; tail -c 400 $out/bench_reference.json; ls -la $out
