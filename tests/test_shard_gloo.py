"""Multi-process sharding on CPU: world_size 2, gloo backend, the test-only emulation library
standing in for the GPU.  Checks that the shards are disjoint and complete and that the gathered rows are
identical to a single-process run."""
import os
import socket

import numpy as np
import pytest

import parity_common as pc
from amaranth_twstft_b200 import shard, synth
from dist_gather import gather_results
from amaranth_twstft_b200.api import Correlator

pytestmark = pytest.mark.skipif(not os.path.exists(pc.EMU_LIB), reason="emulation library not built")
N, NREC = 10000, 7


def test_record_ranges_partition_the_day():
    for nrec in (0, 1, 7, 86400):
        for world in (1, 2, 4, 8):
            r = [shard.record_range(nrec, g, world) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == nrec
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def _records():
    bits, code = pc.make_code(N)
    return bits, np.stack([synth.record_pair(code, 100 + i)[0] for i in range(NREC)])


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bits, recs = _records()
    with Correlator(nsamples=N, max_batch=2, lib=pc.EMU_LIB) as cor:
        cor.set_code(bits)
        lo, local = shard.process_shard(cor, lambda a, b: recs[a:b], NREC, rank, world)
    full = gather_results(local, NREC, 2, rank, world)
    if rank == 0:
        q.put(full)
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_gloo_equal_single_process():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    full = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    bits, recs = _records()
    with Correlator(nsamples=N, max_batch=3, lib=pc.EMU_LIB) as cor:
        cor.set_code(bits)
        ref = cor.process(recs)
    assert full.shape == ref.shape
    for nm in ref.dtype.names:
        assert np.array_equal(full[nm], ref[nm]), nm
