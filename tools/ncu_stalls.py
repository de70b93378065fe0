#!/usr/bin/env python3
"""Key throughput / stall metrics of every launch in an ncu report:  python tools/ncu_stalls.py rep.ncu-rep"""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("====", r[hdr.index("Kernel Name")][:100], r[hdr.index("Grid Size")], r[hdr.index("Block Size")])
    for k in keys:
        if k in hdr:
            print(f"  {k} [{units[hdr.index(k)]}] = {r[hdr.index(k)]}")
    st = []
    for i, h in enumerate(hdr):
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and h.endswith("_not_issued"):
            try: st.append((float(r[i]), h[len("smsp__pcsamp_warps_issue_stalled_"):-len("_not_issued")]))
            except ValueError: pass
    tot = float(r[hdr.index("smsp__pcsamp_sample_count")]) if "smsp__pcsamp_sample_count" in hdr else 1
    print("  no-issue samples by reason (% of all samples): " + ", ".join(f"{n} {100*v/tot:.1f}" for v, n in sorted(st, reverse=True)[:8]))
