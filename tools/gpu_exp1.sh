#!/bin/bash
# experiment batch 1: column-pass L2 prefetch distance, dynamic row kernels, chunk sizes, records-per-step sweep
mkdir -p gpurun_out; rm -f gpurun_out/kbench.jsonl gpurun_out/kbench.err
run() { echo "{\"tag\": \"$1\"}" >> gpurun_out/kbench.jsonl; shift; timeout 300 "$@" >> gpurun_out/kbench.jsonl 2>> gpurun_out/kbench.err; }
V=tools/_variants
run d2 tools/kbench $V/libtwstft_d2.so 16 8 1 4
for pf in 1 2 3; do TWSTFT_COL_PF=$pf run d2_pf$pf tools/kbench $V/libtwstft_d2.so 16 8 1 4; done
for v in rd8 rd4c4 rd16c16; do run $v tools/kbench $V/libtwstft_$v.so 16 8 1 4; done
run d2_again tools/kbench $V/libtwstft_d2.so 16 8 1 4
for R in 2 4 8 32; do run d2_R$R tools/kbench $V/libtwstft_d2.so $R 8 1 4; done
run d2_full tools/kbench $V/libtwstft_d2.so 16 6 0 4
run rd8_full tools/kbench $V/libtwstft_rd8.so 16 6 0 4
tail -3 gpurun_out/kbench.err
