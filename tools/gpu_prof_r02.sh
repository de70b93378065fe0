#!/bin/bash
# Round-2 evidence run on the GPU box: bench.py (the numbers of record), the ncu launch list of the bench command,
# and one ncu --set full capture of one pass of the hot path (tools/prof_run.py).
mkdir -p gpurun_out
TAG=${1:-r02_final}
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_${TAG}_reference.json 2>> gpurun_out/bench_$TAG.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 640 --csv --log-file gpurun_out/launches_bench_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-day > gpurun_out/bench_under_ncu_$TAG.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -s 18 -c 15 -f -o gpurun_out/prof_full_$TAG \
    python tools/prof_run.py 4 1 > gpurun_out/prof_full_$TAG.log 2>&1
tail -2 gpurun_out/prof_full_$TAG.log
python tools/show_bench.py gpurun_out/bench_$TAG.json | head -4 | cut -c1-600
cat gpurun_out/bench_${TAG}_reference.json | cut -c1-400
