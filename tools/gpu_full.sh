#!/bin/bash
# full round check: gpu tests, smoke, bench, ncu launch list, ncu --set full of one hot-path pass
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -o gpurun_out/prof_full -f python tools/prof_run.py 4 1 > gpurun_out/prof_full.log 2>&1
tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -3 gpurun_out/bench.err
