#!/usr/bin/env python3
"""Tiny driver for ncu: one warm-up pass then `passes` passes of the hot path over R resident records."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from amaranth_twstft_b200.api import Correlator
from amaranth_twstft_b200 import synth

R = int(sys.argv[1]) if len(sys.argv) > 1 else 2
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
N = 5000000
fast = os.environ.get('TWSTFT_PROF_FAST', '1') != '0'
cor = Correlator(nsamples=N, max_batch=R, fast_peak=fast)   # the bench's default peak path
bits = cor.gen_code(22, 3)
code = 2.0 * np.repeat(bits.astype(np.float64), 2) - 1.0
rec = synth.record_pair(code, 20261017)[0]
host = np.concatenate([rec] * R)
dev = cor.to_device(host)
for _ in range(1 + passes):
    res = cor.process_device(dev, R)
print(res["indice"][0], cor.kernel_launches)
