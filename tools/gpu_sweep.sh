#!/bin/bash
# H2D yardstick + records-per-step sweep of the bench
mkdir -p gpurun_out
./tools/ubench/h2d | tee gpurun_out/h2d.txt
for r in 4 8 16 32; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --records-per-step $r > gpurun_out/bench_r$r.log 2> gpurun_out/bench_r$r.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r$r.log').read().strip().splitlines()[-1])
print('R=$r value %.1f e2e %.1f clocks %s'%(d['value'],d['e2e']['value'],d['clocks']), ' '.join('%s=%.3f'%(k[:14],v['avg_ms']) for k,v in d['kernels'].items() if v['avg_ms']>0.05))
PY
done
