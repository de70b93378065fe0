#!/usr/bin/env python
"""Static SASS instruction mix per kernel of libtwstft.so (cuobjdump -sass), to compare builds without a GPU."""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "amaranth_twstft_b200/libtwstft.so"
pat = sys.argv[2] if len(sys.argv) > 2 else "KRowCorr|KRow|KCol"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
fn, mix = None, {}
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        mix[fn] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        mix[fn][m.group(1)] += 1
for fn, c in mix.items():
    if not re.search(pat, fn):
        continue
    tot = sum(c.values())
    print(f"{tot:6d} {fn[:110]}")
    print("       " + ", ".join(f"{k} {v}" for k, v in c.most_common(12)))
