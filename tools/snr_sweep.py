#!/usr/bin/env python3
"""Parity and fast-path certificate statistics over input SNR (run on the GPU box):
  python tools/snr_sweep.py [out.json]
N = 5e6 (variant G, two-channel records) and N = 2e5 (variant K windows), SNR -25..-5 dB, 16 seeds per cell, with the
fast peak path; every channel-1 row is compared with the oracle.  The table goes into DESIGN.md."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import parity_common as pc   # noqa: E402

out = {}
for kind, n in (("G", 5000000), ("K", 200000), ("G", 200000)):
    out[f"{kind}_{n}"] = pc.run_snr_sweep(pc.CUDA_LIB, n, kind=kind, snrs=(-25.0, -20.0, -15.0, -10.0, -5.0), strict=False)
    for c in out[f"{kind}_{n}"]:
        print(kind, n, json.dumps(c), flush=True)
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
