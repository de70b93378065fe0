#!/usr/bin/env python3
"""Hot spots of one kernel from `ncu --page source --print-source sass --csv`:
   python tools/ncu_sass_hot.py sass.csv  -> instruction mix, stall totals, top stalled instructions, shared-memory excess."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; body = [r for r in rows[2:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except ValueError: return 0.0
tot_s = sum(f(r, "# Samples") for r in body); tot_i = sum(f(r, "Instructions Executed") for r in body)
print(f"instructions {tot_i:.0f}  samples {tot_s:.0f}  sass lines {len(body)}")
mix = collections.Counter()
for r in body:
    op = r[ix["Source"]].split()[0]
    if op.startswith("@"): op = r[ix["Source"]].split()[1]
    mix[op.split(".")[0]] += f(r, "Instructions Executed")
print("mix: " + ", ".join(f"{k} {100*v/tot_i:.1f}%" for k, v in mix.most_common(22)))
st = {h: sum(f(r, h) for r in body) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
print("stalls(all samples): " + ", ".join(f"{k[6:]} {100*v/tot_s:.1f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:10]))
ex = sum(f(r, "L1 Wavefronts Shared Excessive") for r in body); wv = sum(f(r, "L1 Wavefronts Shared") for r in body)
print(f"shared wavefronts {wv:.0f} excessive {ex:.0f}")
print("-- top shared excess")
for r in sorted(body, key=lambda r: -f(r, "L1 Wavefronts Shared Excessive"))[:12]:
    print(f"  {r[ix['Address']][-5:]} ex {f(r,'L1 Wavefronts Shared Excessive'):.0f} of {f(r,'L1 Wavefronts Shared'):.0f}  {r[ix['Source']][:70]}")
print("-- top sampled instructions")
for r in sorted(body, key=lambda r: -f(r, "# Samples"))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    why = sorted(((f(r, h), h[6:]) for h in hdr if h.startswith("stall_") and "Not Issued" not in h), reverse=True)[:2]
    print(f"  {r[ix['Address']][-5:]} {100*f(r,'# Samples')/tot_s:.2f}%  {r[ix['Source']][:60]:60s} {why[0][1]} {why[0][0]:.0f} {why[1][1]} {why[1][0]:.0f}")
