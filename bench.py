#!/usr/bin/env python3
"""Benchmark of the per-second PRN correlator hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A "step" is one pass of the hot path (carrier estimate -> derotate -> correlate x3 -> peak /
fine delay / phase / SNR) over one batch of RECORDS_PER_STEP synthetic two-channel 1-s records
(5 MS/s, 2.5 Mchip/s BPSK, 22-bit taps-3 LFSR code: BASELINE.json configs[1]).
  value  records/s over all GPUs with the records already resident in HBM (twstft_process_device)
  e2e    the same metric through twstft_process with HOST (pinned) records, H2D + D2H in the timed region
  roofline   dominant kernel: algorithmic bytes per launch / CUDA-event duration vs the measured HBM peak
  cpu_baseline  the numpy oracle (a port of godual_ranging.m) on one host core, one record
  day    BASELINE configs[4]: 86 400 records drawn cyclically from a pool of DAY_POOL distinct records, sharded by
         record range over the ranks, once from HBM-resident pools and once streamed from pinned host
         memory; clocks and power sampled over the whole region; 32 random rows checked against the oracle
--impl reference times the oracle port on all host cores (the reference's Octave/C++ cannot be built here).
The synthetic records are generated on the device (twstft_synth_records: SURVEY 8d recipe, Philox noise).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N = 5000000
FS = 5e6
CODE = (22, 3, 2500000)          # bitlen, taps, noiselen (experiments/221207_twoway_codes/codes/noiselen2500000_bitlen22_taps03)
RECORDS_PER_STEP = 32
DAY_RECORDS = 86400
DAY_POOL = 64                    # distinct records per rank the day is drawn from (record r = pool[r % DAY_POOL])
SEED0 = 20261017
METRIC = "two-channel 1-s 5 MS/s records correlated per second"
UNIT = "records/s"
SURVEY_BYTES_PER_RECORD = 1.4e9  # 140*N bytes per channel, two channels
# ALGORITHMIC bytes per lane (one channel of one record) and launch, in units of N bytes: SURVEY.md section 8d's
# budget, item by item (mean 4N; x^p FFT 12N + 8N; derotate+FFT pass-1 12N; pass-2 32N + IFFT pass-1 32N = the work
# of row_corr; IFFT pass-2 24N; the 16N of the SNR dot product have no kernel here: spectral identity).
ALG_BYTES_N = {"mean": 4, "col_fwd_raw_pow": 12, "col_fwd_raw": 12, "row_fwd_argmax": 8, "row_corr": 64, "col_inv_argmax": 24}
# bytes this design actually has to MOVE per lane and launch: row_corr keeps the product
# spectrum and ffty on chip (tensor memory) and shares the code row through L2, 32N instead of 64N
MOVED_BYTES_N = {"mean": 4, "col_fwd_raw_pow": 12, "col_fwd_raw": 12, "row_fwd_argmax": 8, "row_corr": 32, "col_inv_argmax": 24}
# fast peak path (k_fast.cuh, default): one inverse transform; row_corr reads A 8N, writes B 8N and P 8N; the column
# pass reads 8N and does the job SURVEY budgets at 24N; the exact lags around the maximum read P once more (8N)
ALG_BYTES_N_FAST = {"mean": 4, "col_fwd_raw_pow": 12, "col_fwd_raw": 12, "row_fwd_argmax": 8, "row_corr": 64, "col_inv_argmax": 24}
MOVED_BYTES_N_FAST = {"mean": 4, "col_fwd_raw_pow": 12, "col_fwd_raw": 12, "row_fwd_argmax": 8, "row_corr": 24, "col_inv_argmax": 8,
                      "fast_lag_p": 8}
LAUNCHES_PER_LANE = {}     # every kernel of the table runs once per lane (col_fwd_raw_pow: x^2 estimator pass, col_fwd_raw: derotated pass)
ALG_TOTAL_N = sum(v * LAUNCHES_PER_LANE.get(k, 1) for k, v in MOVED_BYTES_N.items())   # 92
ALG_TOTAL_N_FAST = sum(v * LAUNCHES_PER_LANE.get(k, 1) for k, v in MOVED_BYTES_N_FAST.items())   # 76


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """SM clock and clock-event (throttle) reasons sampled through NVML every millisecond by a thread, ONLY between
    begin() and end() -- i.e. during the timed region (B200_PROFILING.md clocks line; nvidia-smi -lms cannot sample a
    region that lasts tens of milliseconds).  Falls back to one nvidia-smi query if NVML is unavailable."""

    def __init__(self, index):
        self.index, self.h, self.nv = index, None, None
        self.sm, self.pw, self.reasons, self.smax = [], [], set(), None
        self.on = threading.Event()
        self.quit = threading.Event()
        self.thread = None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            idx = self.index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            if vis and all(t.strip().isdigit() for t in vis.split(",")):
                idx = int(vis.split(",")[self.index])
            self.h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.smax = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.nv = nv
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception:
            self.nv = None

    def _sample(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            self.pw.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
        except Exception:
            pass
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        for name, bit in (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown),
                          ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                          ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown),
                          ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap)):
            if r & bit:
                self.reasons.add(name)

    def _run(self):
        while not self.quit.is_set():
            if self.on.wait(0.05):
                try:
                    self._sample()
                except Exception:
                    pass
                time.sleep(0.001)

    def begin(self):
        self.on.set()

    def end(self):
        self.on.clear()

    def stop(self):
        self.quit.set()
        if self.nv is None:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=20).stdout
                f = [float(x) for x in out.strip().split(",")]
                return {"sm_mhz": f[0], "sm_max_mhz": f[1], "reasons": [], "samples": 1,
                        "how": "one nvidia-smi query after the timed region (NVML unavailable)"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"], "samples": 0}
        self.thread.join(timeout=1.0)
        return self.summary()

    def summary(self, since=0):
        sm, pw = self.sm[since:], self.pw[since:]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": float(min(sm)) if sm else None,
                "sm_max_mhz": self.smax, "power_w_median": float(np.median(pw)) if pw else None,
                "power_w_max": float(max(pw)) if pw else None,
                "reasons": sorted(self.reasons), "samples": len(sm),
                "how": "NVML, 1 ms period, sampled only inside the timed regions"}


def bind_to_gpu_numa(index):
    """Pin this process to the CPUs NVML reports as local to GPU `index`, before the pinned staging buffers are
    allocated: with one rank per GPU every rank then streams its 40 MB/record from its own socket's memory."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = int(vis.split(",")[index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else index
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        words = (os.cpu_count() + 63) // 64
        mask = nv.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"bound to {len(cpus)} GPU-local cpus"
    except Exception as e:      # affinity is an optimisation only
        return f"unbound ({type(e).__name__})"
    return "unbound"


def oracle_record(raw, code):
    from oracle import correlator as oc
    freq = oc.freq_axis_G(FS, N)
    k = oc.window(freq, -20000.0, 20000.0)
    return oc.record_G(raw, k, freq, oc.fcode_G(code), code, fs=FS)


def _json_default(o):
    return o.item() if hasattr(o, "item") else str(o)


_REF = {}


def _ref_prepare(args):
    """Untimed: every worker builds its record, code spectrum and axis once."""
    seed, = args
    from oracle import correlator as oc
    from oracle import prn
    from amaranth_twstft_b200 import synth
    code = oc.code_bpsk(prn.lfsr_bits(*CODE))
    raw, _ = synth.record_pair(code, seed, fs=FS)
    freq = oc.freq_axis_G(FS, N)
    _REF.update(code=code, raw=raw, freq=freq, k=oc.window(freq, -20000.0, 20000.0), fcode=oc.fcode_G(code))
    return os.getpid()


def _ref_worker(_):
    """Timed: ONE channel-correlation (half a record) through the oracle port of godual_ranging.m."""
    from oracle import correlator as oc
    t0 = time.perf_counter()
    oc.record_G(_REF["raw"], _REF["k"], _REF["freq"], _REF["fcode"], _REF["code"], fs=FS, channels=(1,))
    return time.perf_counter() - t0


def run_reference(args):
    """Reference arm: the oracle port of godual_ranging.m on all usable host cores (Octave and the
    C++/FFTW port cannot be built in this image).  One step = `workers` channel-correlations in
    parallel (half a record each), sized so that K+W steps end within a few minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    workers = max(1, min(os.cpu_count() or 1, 32))
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(workers) as pool:
        pool.map(_ref_prepare, [(SEED0 + i,) for i in range(workers)], chunksize=1)     # inputs resident before timing
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            pool.map(_ref_worker, range(workers), chunksize=1)
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = 0.5 * workers * len(times) / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: two-channel 1-s 5 MS/s BPSK records, 22-bit taps-3 code, x3 interpolation",
                       "records_per_step": 0.5 * workers},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": "port",
                             "sample": f"{workers} single-channel correlations (half records) per step, numpy oracle "
                                       "port of processing/Octave/godual_ranging.m, one process per core"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line, default=_json_default), flush=True)


def _spot_worker(args):
    raw, chan = args
    from oracle import correlator as oc
    from oracle import prn
    code = oc.code_bpsk(prn.lfsr_bits(*CODE))
    freq = oc.freq_axis_G(FS, N)
    o = oc.record_G(raw, oc.window(freq, -20000.0, 20000.0), freq, oc.fcode_G(code), code, fs=FS, channels=(chan,))[chan]
    return {k: (complex(o[k]) if k == "xval" else float(o[k])) for k in ("indice", "correction", "xval", "SNRr", "SNRi", "df")}


def run_day(cor, args, rank, world, dev, host, P, rec_bytes, code, sampler, barrier, max_over_ranks):
    """BASELINE configs[4]: a day of 86 400 two-channel records, record r drawn from pool slot r % P, the rank's share
    [rank*D/world, (rank+1)*D/world); no collective on the data path.  Resident: pool in HBM.  Streamed:
    every record copied from pinned host memory (double-buffered inside twstft_process).  Device time = CUDA events on
    the plan's stream, max over ranks; the host wall clock of the same region is reported beside it."""
    out = {"records": args.day_records, "pool_records_per_rank": P, "sharding": "record ranges, one rank per GPU, "
           "no collective; rows gathered on the host", "legs": {}}
    rows_res = None
    for leg, nday, src, on_dev in (("resident", args.day_records, dev, True),
                                   ("streamed", min(args.day_streamed_records, args.day_records), host, False)):
        if nday <= 0:
            continue
        lo, hi = nday * rank // world, nday * (rank + 1) // world
        cor.process_cycled(src, P, lo, min(2 * int(cor.cfg.max_batch), hi - lo), on_device=on_dev)    # warm-up
        cor.synchronize()
        f0 = cor.fast_counters
        barrier()
        mark = len(sampler.sm)
        sampler.begin()
        t0 = time.perf_counter()
        cor.timer_start()
        rows = cor.process_cycled(src, P, lo, hi - lo, on_device=on_dev)
        ms = cor.timer_stop()
        wall = time.perf_counter() - t0
        sampler.end()
        barrier()
        ms = max_over_ranks(ms)
        wall = max_over_ranks(wall)
        f1 = cor.fast_counters
        entry = {"records": nday, "seconds_device": ms * 1e-3, "seconds_wall": wall, "records_per_s": nday / (ms * 1e-3),
                 "rank0_fast_peak": {"records": f1[0] - f0[0], "rerun_in_full": f1[1] - f0[1]},
                 "clocks": sampler.summary(mark)}
        if nday < args.day_records:
            entry["seconds_for_a_day_extrapolated"] = ms * 1e-3 * args.day_records / nday
        out["legs"][leg] = entry
        if leg == "resident":
            rows_res, lo_res = rows, lo
        else:   # the two legs draw the same records: rows must agree bit for bit wherever both ran
            m = min(len(rows), len(rows_res)) if lo == lo_res else 0
            entry["rows_equal_resident_leg"] = bool(m and np.array_equal(rows[:m], rows_res[:m]))
    if rank == 0 and rows_res is not None and not args.no_cpu_baseline:
        # parity spot check: 32 random rows of rank 0's share against the oracle, on the host copy of the pool record
        import multiprocessing as mp
        rng = np.random.default_rng(SEED0)
        picks = [(int(rng.integers(0, len(rows_res))), int(rng.integers(1, 3))) for _ in range(32)]
        per = rec_bytes // 2
        jobs = [(np.array(host[((lo_res + r) % P) * per:((lo_res + r) % P + 1) * per]), ch) for r, ch in picks]
        t0 = time.perf_counter()
        with mp.get_context("fork").Pool(min(16, os.cpu_count() or 1)) as pool:
            refs = pool.map(_spot_worker, jobs, chunksize=1)
        worst = {"index_mismatches": 0, "fine_delay_samples": 0.0, "phase_rad": 0.0, "snr_db": 0.0, "df_mismatches": 0}
        for (r, ch), o in zip(picks, refs):
            g = rows_res[r, ch - 1]
            worst["index_mismatches"] += int(g["indice"]) != int(o["indice"])
            worst["df_mismatches"] += g["df"] != o["df"]
            worst["fine_delay_samples"] = max(worst["fine_delay_samples"], abs(g["correction"] - o["correction"]) / 3.0)
            worst["phase_rad"] = max(worst["phase_rad"], abs(np.angle(complex(g["xval_re"], g["xval_im"]) * np.conj(o["xval"]))))
            worst["snr_db"] = max(worst["snr_db"], abs(10 * np.log10(g["SNRr"] + g["SNRi"]) - 10 * np.log10(o["SNRr"] + o["SNRi"])))
        ok = (worst["index_mismatches"] == 0 and worst["df_mismatches"] == 0 and worst["fine_delay_samples"] < 1e-4
              and worst["phase_rad"] < 1e-4 and worst["snr_db"] < 0.01)
        out["parity_spot_check"] = {"rows": len(picks), "ok": bool(ok), "worst": worst,
                                    "tolerances": {"index": "exact", "fine_delay_samples": 1e-4, "phase_rad": 1e-4, "snr_db": 0.01},
                                    "oracle_seconds": time.perf_counter() - t0}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--records-per-step", type=int, default=RECORDS_PER_STEP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-day", action="store_true", help="skip the 86 400-record day leg")
    ap.add_argument("--day-records", type=int, default=DAY_RECORDS)
    ap.add_argument("--day-streamed-records", type=int, default=DAY_RECORDS,
                    help="records of the streamed (host-fed) day leg; the resident leg always runs --day-records")
    ap.add_argument("--no-fast-peak", action="store_true",
                    help="three-polyphase inverse for every record instead of the certified one-transform peak path")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    from amaranth_twstft_b200.api import Correlator
    R = args.records_per_step
    numa = bind_to_gpu_numa(local)
    fast = not args.no_fast_peak
    alg_n = ALG_BYTES_N_FAST if fast else ALG_BYTES_N
    moved_n = MOVED_BYTES_N_FAST if fast else MOVED_BYTES_N
    moved_total_n = ALG_TOTAL_N_FAST if fast else ALG_TOTAL_N
    cor = Correlator(nsamples=N, fs=FS, max_batch=R, device=local, fast_peak=fast)
    assert cor.backend == "cuda"
    bits = cor.gen_code(CODE[0], CODE[1])
    code = 2.0 * np.repeat(bits.astype(np.float64), 2) - 1.0
    # resident pool: P >= R distinct records per rank, generated on the device (R*40 MB >> 126 MB L2 for R >= 4);
    # a pinned host copy of it feeds the end-to-end and streamed legs
    P = max(R, DAY_POOL if not args.no_day else R)
    rec_bytes = 4 * N * 2                          # [I1 Q1 I2 Q2] int16 per sample instant: 40 MB per record
    dev = cor.synth_records(SEED0 + rank * 100000, P)
    pinned = cor.pinned_empty(P * rec_bytes)
    host = pinned.view(np.int16)
    cor._check(cor.lib.twstft_copy_to_host(cor._plan, host.ctypes.data, dev, P * rec_bytes), "twstft_copy_to_host")
    recs = [host[:rec_bytes // 2]]                 # record 0 of this rank, for the CPU baseline leg

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        res = cor.process_device(dev, R)
    cor.synchronize()
    barrier()
    l0 = cor.kernel_launches
    sampler.begin()
    cor.timer_start()
    for _ in range(args.steps):
        res = cor.process_device(dev, R)
    ms = cor.timer_stop()             # records the stop event and synchronises the plan's stream
    sampler.end()
    launches = cor.kernel_launches - l0
    barrier()
    ms = max_over_ranks(ms)
    value = world * R * args.steps / (ms * 1e-3)

    # end to end: host (pinned) records in, results out, copies inside the timed region
    host_r = host[:R * rec_bytes // 2]
    for _ in range(2):
        cor.process(host_r)
    barrier()
    sampler.begin()
    cor.timer_start()
    e2e_steps = max(3, args.steps // 2)
    for _ in range(e2e_steps):
        res_e2e = cor.process(host_r)
    ms_e2e = max_over_ranks(cor.timer_stop())
    sampler.end()
    e2e_value = world * R * e2e_steps / (ms_e2e * 1e-3)
    barrier()

    # host->device ceiling of THIS box at THIS GPU count: the same pinned records copied with nothing else running,
    # all ranks at once (what bounds the end-to-end rate: 40 MB per record over PCIe / host memory)
    scratch = C.c_void_p()
    cor._check(cor.lib.twstft_device_alloc(cor._plan, R * rec_bytes, C.byref(scratch)), "twstft_device_alloc")
    for rep in range(2):
        barrier()
        cor.timer_start()
        for _ in range(4 if rep else 1):
            cor.lib.twstft_copy_to_device(cor._plan, scratch, host_r.ctypes.data, R * rec_bytes)
        ms_h2d = max_over_ranks(cor.timer_stop())
    cor.device_free(scratch)
    h2d_gbs = world * 4 * R * rec_bytes / (ms_h2d * 1e-3) / 1e9
    barrier()

    day = None
    if not args.no_day:
        day = run_day(cor, args, rank, world, dev, host, P, rec_bytes, code, sampler, barrier, max_over_ranks)
    clocks = sampler.stop()

    # per-kernel durations (CUDA events around every launch, two extra steps)
    cor.profile(True)
    for _ in range(2):
        cor.process_device(dev, R)
    cor.profile(False)
    prof = cor.profile_read()
    peak, peak_kind = measured_peak()
    kernels = {}
    for name, (tot_ms, cnt) in prof.items():
        avg = tot_ms / max(cnt, 1)
        ent = {"avg_ms": avg, "launches": cnt, "share": None}
        if name in alg_n or name in moved_n:
            # bytes the kernel has to move for the algorithm it executes -> achieved GB/s
            b = moved_n.get(name, 0) * N * 2 * R
            ent["moved_bytes"] = b
            ent["gbs"] = b / (avg * 1e-3) / 1e9
            ent["frac"] = ent["gbs"] / peak
            # the share of SURVEY 8d's 140N budget that this kernel's job stands for (fusion / the fast peak path
            # remove traffic the budget counts, so this can exceed the bytes moved)
            sb = alg_n.get(name, 0) * N * 2 * R
            ent["survey_budget_bytes"] = sb
            ent["frac_survey_budget"] = sb / (avg * 1e-3) / 1e9 / peak
        kernels[name] = ent
    tot = sum(v[0] for v in prof.values()) or 1.0
    for name in kernels:
        kernels[name]["share"] = prof[name][0] / tot
    dom = max((k for k in kernels if kernels[k].get("moved_bytes")), key=lambda k: prof[k][0], default=None)
    roofline = None
    if dom:
        traffic = None          # dram bytes per launch from the committed ncu --set full capture (profiles/)
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            traffic = tj["traffic_bytes_per_lane"][dom] * 2 * R
        except Exception:
            pass
        roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["gbs"], "peak": peak, "unit": "GB/s",
                    "frac": kernels[dom]["frac"], "traffic": traffic, "peak_kind": peak_kind,
                    "share_of_step": kernels[dom]["share"],
                    "moved_bytes_per_launch": kernels[dom]["moved_bytes"],
                    "survey_budget_bytes_per_launch": kernels[dom]["survey_budget_bytes"],
                    "frac_survey_budget": kernels[dom]["frac_survey_budget"],
                    "pipeline_survey_formula": {"achieved": value / world * SURVEY_BYTES_PER_RECORD / 1e9,
                                                "frac": value / world * SURVEY_BYTES_PER_RECORD / 1e9 / peak,
                                                "what": "SURVEY 8d: records/s per GPU x 1.4e9 B / peak"},
                    "note": "achieved = bytes this kernel must move for the algorithm it executes / CUDA-event "
                            "duration (ncu traffic agrees).  SURVEY 8d budgets more for the same job (product spectrum "
                            "and ffty through HBM, three polyphase inverses): frac_survey_budget and the pipeline "
                            "formula count that budget.  The transforms are FP32-pipe bound (profiles/), not HBM bound"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: two-channel 1-s 5 MS/s BPSK records, 22-bit taps-3 code, x3 interpolation",
                       "records_per_step": R, "samples_per_record": N,
                       "peak_path": "one inverse transform + exact local lags + certificate (k_fast.cuh), uncertified "
                                    "records re-run in full" if fast else "three-polyphase inverse", "l2": f"inputs larger than L2: {R} distinct "
                       f"records = {R * rec_bytes / 1e6:.0f} MB raw, {R * 2 * 200:.0f} MB intermediates per step", "host_numa": numa},
            "clocks": clocks, "gpu_launches": launches,
            "fast_peak": dict(zip(("records", "rerun_in_full"), cor.fast_counters)) if fast else None,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": R * rec_bytes,
                    "d2h_bytes_per_step": int(res_e2e.nbytes),
                    "h2d_ceiling_gbs": h2d_gbs, "h2d_achieved_gbs": e2e_value * rec_bytes / 1e9,
                    "frac_of_h2d_ceiling": e2e_value * rec_bytes / 1e9 / h2d_gbs,
                    "note": "h2d_ceiling = the same pinned records copied host->device with nothing else running, "
                            "all ranks at once on this box: the bound of the end-to-end rate"},
            "day": day,
            "roofline": roofline, "kernels": kernels,
            "pipeline": {"moved_bytes_per_record": moved_total_n * N * 2,
                         "frac_of_hbm_peak_bytes_moved": value / world * moved_total_n * N * 2 / 1e9 / peak,
                         "survey_bytes_per_record": SURVEY_BYTES_PER_RECORD,
                         "frac_of_hbm_peak_survey_bytes": value / world * SURVEY_BYTES_PER_RECORD / 1e9 / peak}}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        t0 = time.perf_counter()
        ref = oracle_record(recs[0], code)
        dt = time.perf_counter() - t0
        ok = all(int(res[0, ch - 1]["indice"]) == ref[ch]["indice"] and
                 abs(res[0, ch - 1]["correction"] - ref[ch]["correction"]) < 3e-4 for ch in (1, 2))
        line["cpu_baseline"] = {"value": 1.0 / dt, "unit": UNIT, "cores": 1, "kind": "port",
                                "sample": "1 two-channel record (2 channel-correlations) through the numpy oracle "
                                          "port of processing/Octave/godual_ranging.m, single process",
                                "parity_with_gpu_on_sample": "ok" if ok else "MISMATCH"}
    if rank == 0:
        print(json.dumps(line, default=_json_default), flush=True)
    cor.device_free(dev)
    cor.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
