#!/bin/bash
# usage: gpurun --timeout 900 -- bash tools/ncu_chain.sh <tag>   -> ncu --set full of the chain kernels (fp4c, sa1) with source correlation
tag=$1; out=gpurun_out/$tag; mkdir -p $out
REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_chain_kernel -c 6 -o $out/chain_fp4c python tools/chain_only.py bf16x3 fp4c > $out/ncu_fp4c.log 2>&1
REPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_chain_kernel -c 2 -o $out/chain_sa1 python tools/chain_only.py bf16x3 sa1 > $out/ncu_sa1.log 2>&1
ls -la $out; tail -3 $out/ncu_fp4c.log
