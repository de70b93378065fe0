#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/kbench.jsonl gpurun_out/kbench.err
run() { echo "{\"tag\": \"$1\"}" >> gpurun_out/kbench.jsonl; shift; timeout 300 "$@" >> gpurun_out/kbench.jsonl 2>> gpurun_out/kbench.err; }
V=tools/_variants
for n in 1 2 3 4 8; do TWSTFT_COL_ITEMS=$n run ci_items$n tools/kbench $V/libtwstft_ci.so 16 6 1 4; done
TWSTFT_COL_ITEMS=2 TWSTFT_COL_PF=1 run ci_items2_pf1 tools/kbench $V/libtwstft_ci.so 16 6 1 4
tail -3 gpurun_out/kbench.err
