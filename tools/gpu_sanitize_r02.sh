#!/bin/bash
# compute-sanitizer memcheck + racecheck over the round-2 kernels on small sizes: smoke() (N = 250 000, fast path,
# code row in tensor memory), the pruned estimator row pass at N = 200 000 (8000-point rows), the complex-double seam,
# the device generator, the sharder and the step-halving acquisition.
mkdir -p gpurun_out
cat > /tmp/san2.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import __graft_entry__ as g
import parity_common as pc
g.smoke()
pc.run_variant_g(pc.CUDA_LIB, 200000, seeds=[11], fast=True)
pc.run_variant_g(pc.CUDA_LIB, 20000, seeds=[21])
pc.run_octave_seam(pc.CUDA_LIB, 200000, seeds=(51,))
pc.run_acquire(pc.CUDA_LIB, 10000)
pc.run_snr_sweep(pc.CUDA_LIB, 10000, kind="G", snrs=(-3.0,), seeds_per_cell=2, procs=2)
print("sanitizer workload ok")
PY
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san2.py > gpurun_out/sanitizer_memcheck_r02.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck_r02.txt
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san2.py > gpurun_out/sanitizer_racecheck_r02.txt 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck_r02.txt
tail -4 gpurun_out/sanitizer_memcheck_r02.txt; tail -4 gpurun_out/sanitizer_racecheck_r02.txt
