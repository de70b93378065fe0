#!/usr/bin/env python3
"""Registers / spills per kernel from a `-Xptxas -v` log:  python tools/ptxas_summary.py log [filter ...]"""
import re, subprocess, sys
txt = open(sys.argv[1]).read()
flt = sys.argv[2:]
for b in re.split(r'ptxas info\s+: Compiling entry function', txt)[1:]:
    m = re.match(r" '(\S+)'", b)
    name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
    name = name.replace('twstft::', '').replace('void kernel_entry<', '')
    name = name[:name.find('>(') + 1] if '>(' in name else name
    if flt and not any(f in name for f in flt):
        continue
    regs = re.search(r'Used (\d+) registers', b)
    sp = re.search(r'(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads', b)
    print(f"{name[:100]:100s} regs {regs.group(1):>3s}  stack/spill-st/spill-ld {'/'.join(sp.groups())}")
