#!/bin/bash
# A/B of an experiment knob given as "$1" (env var name) with values 0/1: parity subset on the default, bench on both
mkdir -p gpurun_out
KNOB=${1:-TWSTFT_SPLITBAR}
if [ -x tools/ubench/warp_slots ]; then timeout 120 tools/ubench/warp_slots > gpurun_out/warp_slots.txt 2>&1; cat gpurun_out/warp_slots.txt; fi
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "not lfsr" 2>&1 | tail -3
for cfg in "0" "1"; do
  echo "=== $KNOB=$cfg"
  env $KNOB=$cfg timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_p$cfg.log 2> gpurun_out/bench_p$cfg.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_p$cfg.log').read().strip().splitlines()[-1])
print('value %.1f e2e %.1f'%(d['value'],d['e2e']['value']), ' '.join('%s=%.3f'%(k[:14],v['avg_ms']) for k,v in d['kernels'].items() if v['avg_ms']>0.05))
PY
done
