"""Test helper: host-side gather of per-rank result rows over torch.distributed (gloo on CPU, NCCL on a GPU box).
The product package does not import torch; a one-process-per-GPU launcher brings its own gather like this one."""
import numpy as np

from amaranth_twstft_b200.api import RESULT_DTYPE
from amaranth_twstft_b200.shard import record_range


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
