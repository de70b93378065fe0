#!/bin/bash
# ncu: launch list of the bench command itself, launch list + full-set capture of one pass of the hot path
mkdir -p gpurun_out
R=${1:-4}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
# fast peak path: 15 launches per pass; 3 (code) + 15 (warm-up pass) = 18 skipped
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 18 -c 15 --csv --log-file gpurun_out/launches.csv python tools/prof_run.py $R 1 > gpurun_out/prof_launch.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -s 18 -c 15 -f -o gpurun_out/prof_full python tools/prof_run.py $R 1 > gpurun_out/prof_full.log 2>&1
ls -la gpurun_out/
tail -3 gpurun_out/prof_full.log
