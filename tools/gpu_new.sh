#!/bin/bash
# new-feature check: the K re/ranging entry points and CLI on the GPU
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "k_re or estimate_bins or two_second or claudio or k_tracking or search_df" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_new.log
tail -15 gpurun_out/pytest_new.log
