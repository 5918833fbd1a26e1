#!/bin/bash
# usage: gpurun --timeout 600 -- bash tools/gpu_quick.sh <tag> "<commands...>"   (each command's output goes to gpurun_out/<tag>/)
tag=$1; shift
out=gpurun_out/$tag; mkdir -p $out
i=0
for cmd in "$@"; do
  i=$((i+1))
  echo "== $cmd" >> $out/log.txt
  ( eval "timeout 400 $cmd" ) >> $out/log.txt 2>&1
done
tail -c 6000 $out/log.txt
