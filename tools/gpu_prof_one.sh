#!/bin/bash
# ncu --set full capture of ONE kernel (regex $1) of the hot path, second launch; report -> gpurun_out/prof_$2.ncu-rep
mkdir -p gpurun_out
K=${1:-KRowCorr}; TAG=${2:-one}; R=${3:-4}
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$K -s 1 -c 1 -f -o gpurun_out/prof_$TAG python tools/prof_run.py $R 1 > gpurun_out/prof_$TAG.log 2>&1
tail -2 gpurun_out/prof_$TAG.log
