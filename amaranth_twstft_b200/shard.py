"""Record sharding across the GPUs of one box.

Every 1-s record is independent in variant G (processing/Octave/godual_ranging.m:75-102 keeps no state
between iterations except the append-only result arrays), so a day of records is split by record index:
rank g of G owns records [g*R/G, (g+1)*R/G).  One process per GPU, one plan per process, the code spectrum
replicated per GPU.  No collective touches the data path; the result rows (128 bytes per channel-record)
are gathered on rank 0 at the end through torch.distributed (NCCL on the GPU box, gloo in the CPU tests).
"""
import numpy as np

from .api import RESULT_DTYPE


def record_range(nrecords, rank, world):
    """Contiguous, balanced, disjoint ranges that cover [0, nrecords)."""
    return nrecords * rank // world, nrecords * (rank + 1) // world


def process_shard(cor, read_records, nrecords, rank=0, world=1, batch=None):
    """Run this rank's share.  read_records(lo, hi) -> int16 array holding records lo..hi-1 (host memory).
    Returns (lo, results[hi-lo, nch])."""
    lo, hi = record_range(nrecords, rank, world)
    batch = batch or int(cor.cfg.max_batch)
    out = np.zeros((hi - lo, cor.nch), dtype=RESULT_DTYPE)
    for a in range(lo, hi, batch):
        b = min(hi, a + batch)
        out[a - lo:b - lo] = cor.process(read_records(a, b))
    return lo, out


def gather_results(local, nrecords, nch, rank=0, world=1):
    """Host-side gather of the per-rank result rows; returns the full (nrecords, nch) array on rank 0 and
    None elsewhere.  Uses all_gather on a padded float64 tensor (works with both gloo and nccl)."""
    if world == 1:
        return local
    import torch
    import torch.distributed as dist
    nf = len(RESULT_DTYPE.names)
    cap = max(record_range(nrecords, r, world)[1] - record_range(nrecords, r, world)[0] for r in range(world))
    flat = np.zeros((cap, nch, nf), dtype=np.float64)
    flat[:len(local)] = np.ascontiguousarray(local).view(np.float64).reshape(len(local), nch, nf)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    mine = torch.from_numpy(flat).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    if rank != 0:
        return None
    full = np.zeros((nrecords, nch), dtype=RESULT_DTYPE)
    for r in range(world):
        lo, hi = record_range(nrecords, r, world)
        full[lo:hi] = np.ascontiguousarray(parts[r].cpu().numpy()[:hi - lo]).view(RESULT_DTYPE).reshape(hi - lo, nch)
    return full
