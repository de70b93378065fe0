#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "fast_peak or variant_g or reference_python or host_and_device" 2>&1 | tail -3
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fast.log 2> gpurun_out/bench_fast.err
tail -2 gpurun_out/bench_fast.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_fast.log').read().strip().splitlines()[-1])
print('value %.1f e2e %.1f launches %d fast %s'%(d['value'],d['e2e']['value'],d['gpu_launches'],d.get('fast_peak')))
print(' '.join('%s=%.3f'%(k[:18],v['avg_ms']) for k,v in d['kernels'].items() if v['avg_ms']>0.02))
PY
