#!/usr/bin/env python3
"""Summarise an `ncu --set full` report (per launch) as a markdown table for profiles/.

  python tools/ncu_summary.py gpurun_out/prof_full.ncu-rep [lanes] > profiles/rNN_ncu_full.md

Reads the report with `ncu -i ... --page raw --csv` (B200_PROFILING.md recipe).  `lanes` = channel
correlations per launch of the profiled run (tools/prof_run.py R records -> 2R lanes); it turns the
DRAM traffic into bytes per lane in units of N = 5e6 samples so that it can be compared with the
algorithmic bytes of DESIGN.md.
"""
import csv
import io
import json
import os
import subprocess
import sys

COLS = [
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_rd_MB"),
    ("dram__bytes_write.sum", "dram_wr_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma_pipe_pct"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "smem_KB"),
    ("smsp__inst_executed.sum", "warp_inst"),
    ("lts__t_sector_hit_rate.pct", "l2_hit_pct"),
    ("derived__memory_l1_wavefronts_shared_excessive", "smem_excess_wavefronts"),
]


def short(name):
    name = name.replace("void kernel_entry<", "").replace(">(Params)", "").replace("FFTSpec", "F")
    return name[:60]


def main():
    rep = sys.argv[1]
    lanes = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    n = 5e6
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    out = []
    print(f"# ncu --set full summary of `{os.path.basename(rep)}` ({lanes} lanes per launch, N = 5e6)\n")
    print("One row per launch of one pass of the hot path (tools/prof_run.py).  Times are under the profiler"
          " (cold cache, serialised): use shares, not absolutes.  `traffic_N` = (dram read + write) per lane in units"
          " of N bytes, to be compared with the algorithmic bytes of DESIGN.md.\n")
    names = ["kernel", "grid", "block"] + [c[1] for c in COLS] + ["traffic_N"]
    print("| " + " | ".join(names) + " |")
    print("|" + "---|" * len(names))
    tot = 0.0
    for r in body:
        vals = {}
        for full, shortn in COLS:
            try:
                v = float(r[ix[full]])
                u = units[ix[full]]
                if shortn == "time_us":
                    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
                if shortn.endswith("_MB"):
                    v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                if shortn == "smem_KB":
                    v *= {"byte": 1e-3, "Kbyte": 1.0}.get(u, 1.0)
            except Exception:
                v = float("nan")
            vals[shortn] = v
        traffic = (vals["dram_rd_MB"] + vals["dram_wr_MB"]) * 1e6 / lanes / n
        tot += vals["time_us"]
        cells = [short(r[ix["Kernel Name"]]), r[ix["Grid Size"]], r[ix["Block Size"]]]
        for _, shortn in COLS:
            v = vals[shortn]
            cells.append(f"{v:.0f}" if shortn in ("regs", "warp_inst", "smem_excess_wavefronts") else f"{v:.1f}")
        cells.append(f"{traffic:.1f}")
        print("| " + " | ".join(cells) + " |")
        out.append({"kernel": short(r[ix["Kernel Name"]]), **vals, "traffic_bytes_per_lane": traffic * n})
    print(f"\nTotal profiled time {tot:.0f} us; shares: " +
          ", ".join(f"{o['kernel'].split('<')[0]}{'' if '<' not in o['kernel'] else '<..' + o['kernel'][-8:]} {100 * o['time_us'] / tot:.0f}%" for o in out))
    json.dump(out, open(os.path.splitext(sys.argv[3])[0] + ".json", "w"), indent=1) if len(sys.argv) > 3 else None


if __name__ == "__main__":
    main()
