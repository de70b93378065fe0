#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/kbench.jsonl gpurun_out/kbench.err
run() { echo "{\"tag\": \"$1\"}" >> gpurun_out/kbench.jsonl; shift; timeout 300 "$@" >> gpurun_out/kbench.jsonl 2>> gpurun_out/kbench.err; }
V=tools/_variants
for v in $VARIANTS; do run $v tools/kbench $V/libtwstft_$v.so 16 6 ${FAST:-1} 4; done
tail -3 gpurun_out/kbench.err
