#!/usr/bin/env python3
"""Production (K) shape benchmark: `claudio_aligned lo` (processing/C) on a synthetic single-channel X310 recording
-- 100 kchip code (N = 200 000, 25 windows/s), 2-s blocks, 5 MS/s -- generated on the device and written to a file.
Reports seconds of recorded data processed per second of wall clock, file-to-.mat, and the per-kernel CUDA-event
totals of the two plans (TWSTFT_PROFILE_DUMP).  Reference being replaced: acquisition/goprocess.sh:9-11 ->
claudio_aligned_code_lo_separate.m:114-164 (Octave, "~60 min per 3-min dataset", experiments/221219_twoway/README.md:8-10).

  python tools/bench_k.py [seconds=330] [outdir=/tmp/twstft_k] [snr_db=-10]
"""
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from amaranth_twstft_b200.api import Correlator   # noqa: E402

secs = int(sys.argv[1]) if len(sys.argv) > 1 else 330
outdir = sys.argv[2] if len(sys.argv) > 2 else "/tmp/twstft_k"
snr_db = float(sys.argv[3]) if len(sys.argv) > 3 else -10.0
fs, n = 5e6, 200000
os.makedirs(outdir, exist_ok=True)
data = os.path.join(outdir, "1700000000_1700000330.5_1.bin")
codef = os.path.join(outdir, "code.bin")
t0 = time.perf_counter()
with Correlator(nsamples=n, fs=fs, variant="K", layout="single", max_batch=1) as cor:
    bits = cor.gen_code(17, 9, noiselen=n // 2)
    open(codef, "wb").write(bits.tobytes())
    with open(data, "wb") as f:
        chunk = 10                                   # seconds per generated piece (200 MB)
        for s0 in range(0, secs, chunk):
            m = min(chunk, secs - s0)
            # the code repeats back to back: a piece starting at second s0 continues the same phase
            ptr = cor.synth_stream(4242 + s0, int(m * fs), delay=1234.0, df=4321.7, snr_db=snr_db)
            f.write(cor.from_device(ptr, int(m * fs) * 4).tobytes())
            cor.device_free(ptr)
t_gen = time.perf_counter() - t0
exe = os.path.join(ROOT, "processing", "C", "claudio_aligned")
env = dict(os.environ, TWSTFT_PROFILE_DUMP="1")
runs = []
for rep in range(2):                                 # second run: file in the page cache
    t0 = time.perf_counter()
    out = subprocess.run([exe, "lo", data, codef, "0", outdir], capture_output=True, text=True, env=env)
    dt = time.perf_counter() - t0
    assert out.returncode == 0, out.stderr[-2000:]
    runs.append(dt)
rows = [ln for ln in out.stdout.splitlines() if ln and ln[0].isdigit()]
prof = {}
for ln in out.stderr.splitlines():
    if ln.startswith("twstft_profile"):
        tag, js = ln.split(" ", 2)[1:]
        prof[tag] = json.loads(js)
tot = sum(v[0] for p in prof.values() for v in p.values()) or 1.0
line = {"bench": "claudio_aligned lo (K shape)", "recording_s": secs, "windows": len(rows), "snr_db": snr_db,
        "wall_s_cold": runs[0], "wall_s_warm": runs[1], "seconds_of_data_per_second": secs / runs[1],
        "windows_per_s": len(rows) / runs[1], "file_gb": os.path.getsize(data) / 1e9, "generate_s": t_gen,
        "gpu_kernel_ms_total": tot,
        "gpu_busy_fraction_of_wall": tot * 1e-3 / runs[1],
        "kernels": {f"{tag}:{k}": {"ms": v[0], "launches": v[1], "share": v[0] / tot}
                    for tag, p in prof.items() for k, v in p.items() if v[0] / tot > 0.005},
        "tail": out.stderr.strip().splitlines()[-1]}
print(json.dumps(line))
