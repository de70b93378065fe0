#!/bin/bash
# two-rank check of the sharded bench (one process per GPU, torchrun as the driver launches it) + the reference arm
mkdir -p gpurun_out
P=29517
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) \
  bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_ref_n2.log 2> gpurun_out/bench_ref_n2.err; echo "ref n2 rc=$?"
tail -2 gpurun_out/bench_n2.log | cut -c1-1500; tail -5 gpurun_out/bench_n2.err; tail -1 gpurun_out/bench_ref_n2.log | cut -c1-800
nproc; nvidia-smi topo -m 2>/dev/null | head -8
