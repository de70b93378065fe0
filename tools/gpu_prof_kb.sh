#!/bin/bash
# ncu --set full of one kernel (regex $2) of a kbench run on variant library $1 -> gpurun_out/prof_$3.ncu-rep
mkdir -p gpurun_out
V=$1; K=$2; TAG=$3
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$K" -s 1 -c 1 -f -o gpurun_out/prof_$TAG tools/kbench tools/_variants/libtwstft_$V.so 4 2 1 2 > gpurun_out/prof_$TAG.log 2>&1
tail -2 gpurun_out/prof_$TAG.log | cut -c1-200
