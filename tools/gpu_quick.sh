#!/bin/bash
# quick check: gpu tests (fast subset) + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "not lfsr" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -4 gpurun_out/pytest_gpu.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],'launches',d['gpu_launches'],'clocks',d['clocks'])
for k,v in d['kernels'].items():
    print(f"  {k:18s} avg_ms {v['avg_ms']:.4f} share {v['share']:.3f} frac {v.get('frac',0):.3f}")
print('pipeline',d['pipeline']); print('cpu',d.get('cpu_baseline'))
PY
tail -3 gpurun_out/bench.err
