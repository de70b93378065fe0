#!/bin/bash
# compute-sanitizer memcheck + racecheck on a small record (N = 250 000: same kernels as the 5e6 case), bank timing
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import __graft_entry__ as g
g.smoke()
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san.py > gpurun_out/sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san.py > gpurun_out/sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.txt
tail -4 gpurun_out/sanitizer_memcheck.txt; tail -4 gpurun_out/sanitizer_racecheck.txt
timeout 300 python tools/bench_bank.py > gpurun_out/bench_bank.json 2> gpurun_out/bench_bank.err; cat gpurun_out/bench_bank.json; tail -2 gpurun_out/bench_bank.err
