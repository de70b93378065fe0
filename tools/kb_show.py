#!/usr/bin/env python
"""Print gpurun_out/kbench.jsonl (tools/kbench lines, optionally preceded by {"tag": ...} lines) as a table."""
import json, sys
tag = ""
for l in open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/kbench.jsonl"):
    try:
        d = json.loads(l)
    except Exception:
        continue
    if "tag" in d and "lib" not in d:
        tag = d["tag"]; continue
    k = d["kernels"]
    big = " ".join(f"{n[:12]}={v:.4f}" for n, v in k.items() if v > 0.05)
    rows = d.get("rows")
    h = hash(json.dumps(rows)) & 0xffff if rows is not None else 0
    print(f"{tag:12s} R={d['R']:2d} {d['records_per_s']:7.1f} rec/s  {d['us_per_lane']:6.2f} us/lane fast={d['fast']} rows#{h:04x} | {big}")
