#!/bin/bash
# Round-2 evidence on one B200: parity tests, smoke, bench (both arms), ncu launch list + full capture + source pages, profiles.
out=gpurun_out/final2; mkdir -p $out
( timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | tail -6 ) > $out/pytest.txt
( timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 ) > $out/smoke.txt
timeout 900 python bench.py --steps 32 --warmup 8 > $out/bench.json 2> $out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err
timeout 300 python bench.py --steps 32 --warmup 8 --precision bf16 --no-extras --no-cpu-baseline > $out/bench_bf16.json 2> /dev/null
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $out/launches.csv python tools/one_forward.py > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o $out/fwd python tools/one_forward.py > $out/ncu_full.log 2>&1
timeout 400 python tools/ablate.py 8 > $out/ablate.txt 2>&1
timeout 100 python tools/pcie_probe.py > $out/pcie.txt 2>&1
timeout 200 python tools/tc_profile.py bf16x3 > $out/tc_profile.txt 2>&1
timeout 200 python tools/fps_profile.py > $out/fps_profile.txt 2>&1
timeout 200 python tools/chain_only.py bf16x3 sa1 sa2 sa3 sa4 fp1 fp2 fp3 fp4 fp4c > $out/chain_only.txt 2>&1
timeout 200 python tools/chain_only.py bf16 sa1 sa2 sa3 sa4 fp1 fp2 fp3 fp4 fp4c >> $out/chain_only.txt 2>&1
timeout 200 python bench.py --workload cfg3 --steps 20 2>/dev/null | tail -1 > $out/cfg3.json
timeout 200 python bench.py --workload cfg3 --steps 20 --precision bf16 2>/dev/null | tail -1 > $out/cfg3_bf16.json
timeout 200 python bench.py --workload cfg4 --steps 6 2>/dev/null | tail -1 > $out/cfg4.json
timeout 300 python tools/sweep_cfg5.py > $out/cfg5.txt 2>&1; cp gpurun_out/cfg5_sweep.json $out/cfg5_sweep.json 2>/dev/null
for tool in memcheck racecheck synccheck; do ( timeout 500 compute-sanitizer --tool $tool python tools/sanitize.py 2>&1 | tail -4 ) > $out/sanitizer_$tool.txt; done
cat $out/pytest.txt $out/smoke.txt; tail -c 300 $out/bench.json; echo; tail -c 300 $out/bench_reference.json; ls -la $out
