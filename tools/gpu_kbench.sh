#!/bin/bash
# usage (on the GPU box): tools/gpu_kbench.sh "<variant names>" [R] [steps]; appends JSON lines to gpurun_out/kbench.jsonl
mkdir -p gpurun_out
R=${2:-16}; STEPS=${3:-6}
for v in $1; do
  for fast in 1 0; do
    timeout 300 tools/kbench tools/_variants/libtwstft_$v.so $R $STEPS $fast 4 >> gpurun_out/kbench.jsonl 2>> gpurun_out/kbench.err
  done
done
