#!/usr/bin/env python3
"""Turn one evidence run of tools/gpu_prof_r02.sh into the committed summaries under profiles/:

  python tools/prof_digest.py r02_c

reads  gpurun_out/bench_<tag>.json, gpurun_out/launches_bench_<tag>.csv, gpurun_out/prof_full_<tag>.ncu-rep
writes profiles/<tag>_ncu_launch_shares.md   per-kernel share of the resident step: ncu launch list vs bench.py CUDA events
       profiles/<tag>_ncu_sass_stalls.md     SASS instruction mix + warp-stall samples of the hot kernels (source page)
       profiles/ncu_traffic.json             DRAM bytes per lane and kernel (bench.py's roofline.traffic reads it)
(the --set full table itself comes from tools/ncu_summary.py).
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02_c"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
NPTS = 5e6

NAMES = [  # (regex on the demangled kernel name, bench.py kernel name)
    (r"KMean", "mean"),
    (r"KCol<.*\b16, 0, 0, 2, 8000, 2>", "col_fwd_raw_pow"),
    (r"KCol<.*\b16, 0, 0, 1, 8000, 2>", "col_fwd_raw"),
    (r"KRow<.*, 1, 1>", "row_fwd_argmax"),
    (r"KRowCorr<.*, 1, 1>", "row_corr"),
    (r"KCol<.*\b8, 1, 1, 0, 8000, 1>", "col_inv_argmax"),
    (r"KLagP", "fast_lag_p"),
]


def clean(n):
    return re.sub(r"\(int\)", "", n).replace("twstft::", "")


def bench_name(kname):
    k = clean(kname)
    for rx, b in NAMES:
        if re.search(rx, k):
            return b
    return None


# ---------------------------------------------------------------- launch shares
bench = json.loads(open(os.path.join(G, f"bench_{tag}.json")).read().strip().splitlines()[-1])
lanes = 2 * bench["config"].get("records_per_step", 32)
rows = [r for r in csv.reader(l for l in open(os.path.join(G, f"launches_bench_{tag}.csv")) if l.startswith('"'))]
hdr = rows[0]
ki, gi, vi = hdr.index("Kernel Name"), hdr.index("Grid Size"), hdr.index("Metric Value")
launches = [(clean(r[ki]), r[gi], float(r[vi].replace(",", "")) * 1e-6) for r in rows[1:]]
# the resident steps: runs of launches that start with the mean kernel over the whole batch; the end-to-end leg
# that follows works in one-record chunks (much shorter launches of the same kernels) and is not counted
mean_ms = [ms for k, g, ms in launches if k.startswith("void kernel_entry<KMean>")]
big = max(mean_ms)
acc, cnt, in_step = collections.defaultdict(float), collections.Counter(), False
for k, g, ms in launches:
    if "KMean" in k:
        in_step = ms > 0.5 * big
    if not in_step:
        continue
    b = bench_name(k) or "(small kernels)"
    acc[b] += ms
    cnt[b] += 1
tot = sum(acc.values())
btot = sum(v["avg_ms"] * v["count"] for v in bench["kernels"].values()) if all("count" in v for v in bench["kernels"].values()) else None
out = [f"# Per-kernel share of the resident step: ncu launch list of the bench command vs bench.py's CUDA events ({tag})", "",
       f"`gpurun_out/launches_bench_{tag}.csv` (committed as `profiles/{tag}_ncu_launches_bench.csv`) = `ncu --metrics gpu__time_duration.sum "
       "--clock-control none -c 640 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-day`; counted are the launches of the resident "
       f"steps ({lanes} lanes per launch).  ncu times are serialised and cold-cache: the SHARES must agree, not the absolutes.", "",
       "| kernel | launches | ncu ms/launch | ncu share | bench.py ms/launch | bench.py share |", "|---|---|---|---|---|---|"]
for b in [n for _, n in NAMES] + ["(small kernels)"]:
    if b not in acc:
        continue
    bk = bench["kernels"].get(b)
    if bk:
        out.append(f"| {b} | {cnt[b]} | {acc[b] / cnt[b]:.3f} | {100 * acc[b] / tot:.1f} % | {bk['avg_ms']:.3f} | {100 * bk['share']:.1f} % |")
    else:
        small = sum(v["share"] for k, v in bench["kernels"].items() if k not in [n for _, n in NAMES])
        out.append(f"| {b} | {cnt[b]} | | {100 * acc[b] / tot:.1f} % | | {100 * small:.1f} % |")
open(os.path.join(P, f"{tag}_ncu_launch_shares.md"), "w").write("\n".join(out) + "\n")

# ---------------------------------------------------------------- traffic per lane from the --set full capture
rep = os.path.join(G, f"prof_full_{tag}.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
h, units = rr[0], rr[1]


def val(r, key):
    i = h.index(key)
    x = float(r[i].replace(",", ""))
    u = units[i]
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


prof_lanes = 8
traffic = {}
for r in rr[2:]:
    b = bench_name(r[h.index("Kernel Name")])
    if b and b not in traffic:
        traffic[b] = (val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")) / prof_lanes
tpath = os.path.join(P, "ncu_traffic.json")
tj = json.load(open(tpath)) if os.path.exists(tpath) else {}
tj["source"] = f"profiles/{tag}_ncu_full.md (ncu --set full, tools/prof_run.py 4 records = {prof_lanes} lanes per launch, fast peak path)"
tj["unit"] = "dram bytes (read+write) per lane (one channel of one record) and launch"
tj["traffic_bytes_per_lane"] = traffic
json.dump(tj, open(tpath, "w"), indent=1)

# ---------------------------------------------------------------- SASS mix and stalls per hot kernel
md = [f"# SASS instruction mix and warp-stall samples of the hot kernels ({tag})", "",
      f"From `gpurun_out/prof_full_{tag}.ncu-rep` (`ncu --set full --import-source on`, tools/prof_run.py 4 records = 8 lanes per launch), "
      "source page, SASS view of the first launch of each kernel.  mix = executed warp instructions by opcode (share of the kernel's total; "
      "thread instructions per complex point in brackets); stalls = share of all warp samples.", ""]
for rx, b in NAMES[1:]:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv", "--kernel-name-base", "demangled",
                          "-k", "regex:" + rx.replace("<", ".").replace(">", ".").replace(r"\b", "").replace(", ", ",.*"), "-c", "1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    hd = rows[hi]
    body = [r for r in rows[hi + 1:] if len(r) == len(hd)]
    first, seen = [], set()
    for r in body:                      # the csv may hold the listing twice (two views); keep the first copy
        a = r[hd.index("Address")]
        if a in seen:
            break
        seen.add(a)
        first.append(r)
    body = first
    ix = {k: i for i, k in enumerate(hd)}

    def f(r, k):
        try:
            return float(r[ix[k]])
        except ValueError:
            return 0.0
    tot_s = sum(f(r, "# Samples") for r in body)
    tot_i = sum(f(r, "Instructions Executed") for r in body)
    mix = collections.Counter()
    for r in body:
        t = r[ix["Source"]].split()
        op = t[1] if t[0].startswith("@") else t[0]
        mix[op.split(".")[0]] += f(r, "Instructions Executed")
    per_pt = 32.0 / (prof_lanes * NPTS)
    st = {k: sum(f(r, k) for r in body) for k in hd if k.startswith("stall_") and "Not Issued" not in k}
    md += [f"## {b}", "", f"{len(body)} SASS lines, {tot_i / 1e6:.1f} M warp instructions = {tot_i * per_pt:.1f} thread instructions per complex point, "
           f"{tot_s:.0f} samples", "",
           "mix: " + ", ".join(f"{k} {100 * v / tot_i:.1f}% [{v * per_pt:.2f}]" for k, v in mix.most_common(16)), "",
           "stalls: " + ", ".join(f"{k[6:]} {100 * v / tot_s:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]), "",
           "top sampled instructions:", "", "```"]
    for r in sorted(body, key=lambda r: -f(r, "# Samples"))[:8]:
        why = max(((f(r, k), k[6:]) for k in st), default=(0, ""))
        md.append(f"{100 * f(r, '# Samples') / tot_s:5.2f}%  {why[1]:14s} {r[ix['Source']].strip()[:90]}")
    md += ["```", ""]
    tm = sum(v for k, v in mix.items() if k in ("LDTM", "STTM"))
    if tm:
        nl = sum(1 for r in body if r[ix["Source"]].split()[0].startswith(("LDTM", "STTM")))
        md += [f"tensor-memory instructions in this kernel: {nl} static LDTM/STTM", ""]
open(os.path.join(P, f"{tag}_ncu_sass_stalls.md"), "w").write("\n".join(md))
print("wrote", f"profiles/{tag}_ncu_launch_shares.md", f"profiles/{tag}_ncu_sass_stalls.md", "profiles/ncu_traffic.json")
