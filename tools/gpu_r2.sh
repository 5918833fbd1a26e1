#!/bin/bash
# usage: gpurun --timeout 900 -- bash tools/gpu_r2.sh <tag> "<cmd>" "<cmd>" ...   each command under its own timeout; logs to gpurun_out/<tag>/
tag=$1; shift
out=gpurun_out/$tag; mkdir -p $out
i=0
for cmd in "$@"; do
  i=$((i+1))
  echo "== [$i] $cmd" >> $out/log.txt
  timeout 600 bash -c "$cmd" >> $out/log.txt 2>&1
  echo "== [$i] rc=$?" >> $out/log.txt
done
tail -c 12000 $out/log.txt
