#!/bin/bash
# usage: gpurun --timeout 600 -- bash tools/gpu_quick.sh <tag> "<commands...>"   (each command's output goes to gpurun_out/<tag>/log.txt)
tag=$1; shift
out=gpurun_out/$tag; mkdir -p $out
for cmd in "$@"; do
  echo "== $cmd" >> $out/log.txt
  timeout 400 bash -c "$cmd" >> $out/log.txt 2>&1
done
tail -c 7000 $out/log.txt
