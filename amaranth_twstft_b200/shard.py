"""Record sharding across the GPUs of one box.

Every 1-s record is independent in variant G (processing/Octave/godual_ranging.m:75-102 keeps no state
between iterations except the append-only result arrays), so a day of records is split by record index:
rank g of G owns records [g*R/G, (g+1)*R/G).  One process per GPU, one plan per process, the code spectrum
replicated per GPU.  No collective touches the data path.  Inside ONE process the C library shards by itself
(twstft_shard_*, api.Sharder: one host thread + plan per GPU); these helpers are for the one-process-per-GPU
launch the benchmark contract uses, where every rank calls process_shard on its own range and the launcher gathers
the 128-byte rows.  This module imports neither torch nor NCCL.
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
