"""CPU (gloo, world_size 2 and 3): the multi-GPU host logic -- row-block sharding aligned to partition blocks, the
single end-of-run all-gather and the reassembly order.  The per-rank compute is the oracle
restricted to the rank's rows, so the assembled result must equal the unsharded oracle run (up to BLAS rounding)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import reach_oracle as orc
from reachability_jl_b200 import sharding


def test_shard_blocks_balanced_and_whole():
    sizes = [2] * 503
    seen = []
    for world in (1, 2, 4, 8, 3):
        got = [sharding.shard_blocks(sizes, r, world) for r in range(world)]
        assert got[0][0] == 0 and got[-1][1] == len(sizes)
        for a, b in zip(got, got[1:]):
            assert a[1] == b[0]                    # contiguous, no gaps, no overlap
        rows = [2 * (hi - lo) for lo, hi in got]
        assert max(rows) - min(rows) <= 2          # balanced by rows up to one block
        seen.append(rows)
    # ragged blocks: never split a block, every block owned exactly once
    sizes = [1, 4, 2, 2, 7, 1, 1, 3]
    for world in (2, 3, 5, 8, 16):
        got = [sharding.shard_blocks(sizes, r, world) for r in range(world)]
        owned = [b for lo, hi in got for b in range(lo, hi)]
        assert owned == list(range(len(sizes)))


def test_shard_rows_follow_partition():
    partition = [[1, 2], [3, 4], [5], [6, 7, 8]]
    blocks = [1, 3, 4]
    r0 = sharding.shard_rows(partition, blocks, 0, 2)
    r1 = sharding.shard_rows(partition, blocks, 1, 2)
    assert list(r0) + list(r1) == [0, 1, 4, 5, 6, 7]
    assert list(sharding.shard_rows(partition, blocks, 0, 1)) == [0, 1, 4, 5, 6, 7]


def test_shard_batch_covers_everything():
    for batch, world in ((4096, 8), (10, 3), (2, 4)):
        idx = [i for r in range(world) for i in range(batch)[sharding.shard_batch(batch, r, world)]]
        assert idx == list(range(batch))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = np.random.default_rng(0)
        n, m, N = 14, 2, 25
        A = g.standard_normal((n, n)) / np.sqrt(n) - np.eye(n)
        D = orc.discretize_box(A, orc.Box(g.standard_normal(n), np.abs(g.standard_normal(n)) * 0.1), 0.02,
                               g.standard_normal((n, m)), orc.Box(np.zeros(m), np.ones(m)))
        partition = [[1, 2], [3, 4], [5], [6, 7, 8], [9, 10], [11, 12], [13, 14]]
        blocks = [1, 2, 4, 5, 7]

        def compute(rows):
            C, R = orc.bffpsv18_reach(D, N, rows)
            return C.T, R.T

        C, R = sharding.bffpsv18_sharded(compute, partition, blocks, dist)
        rows_all = [d - 1 for b in blocks for d in partition[b - 1]]
        Cf, Rf = orc.bffpsv18_reach(D, N, rows_all)
        # BLAS may block a row subset differently from the full matrix: equality up to rounding, far inside 1e-10
        ok = (C.shape == Cf.T.shape and np.allclose(C, Cf.T, rtol=1e-13, atol=1e-15) and np.allclose(R, Rf.T, rtol=1e-13, atol=1e-15))
        if rank == 0:
            q.put(bool(ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_row_sharded_flowpipe_equals_unsharded(world):
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get() is True


class _MockEngine:
    """Stands in for backend.ShardedFlowpipe: records the call sequence; every buffer starts as rank + 1."""

    def __init__(self, rank, nb=3):
        self.rank, self.nb, self.log, self.bufs = rank, nb, [], {}

    def nbatches(self):
        return self.nb

    def phase(self, b, ph):
        self.log.append(("phase", b, ph))
        for which in sharding.SHARD_COLLECTIVES[ph]:
            if not (b == 0 and which in (2, 3)):             # the initialisation batch has no box partials
                self.bufs[(b, which)] = torch.full((4,), float(self.rank + 1), dtype=torch.float64)

    def buffer(self, b, which):
        self.log.append(("buffer", b, which))
        return self.bufs.get((b, which))


def _sharded_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        eng = _MockEngine(rank)
        sharding.run_sharded(eng, dist, as_tensor=lambda t: t, sync=lambda: None)
        want = float(sum(range(1, world + 1)))
        ok = all(bool(torch.all(t == want)) for t in eng.bufs.values())
        # order: phases 0..3 of batch 0, then batch 1, ...; the collectives of a phase directly after it
        phases = [e for e in eng.log if e[0] == "phase"]
        ok = ok and phases == [("phase", b, ph) for b in range(3) for ph in range(4)]
        i0 = eng.log.index(("phase", 1, 0))
        ok = ok and eng.log[i0 + 1:i0 + 3] == [("buffer", 1, 0), ("buffer", 1, 1)] and eng.log[i0 + 3] == ("phase", 1, 1)
        if rank == 0:
            q.put(bool(ok))
    finally:
        dist.destroy_process_group()


def test_gloo_run_sharded_places_the_collectives():
    """`run_sharded`: four phases per batch in order, sum all-reduce of the reported buffers right after their phase."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get() is True


class _SingleRankDist:
    """world_size-1 stand-in for torch.distributed (no process group needed): the all-reduce is the identity."""

    class ReduceOp:
        SUM = "sum"

    @staticmethod
    def all_reduce(t, op=None):
        return t


def test_run_sharded_synchronises_around_collectives_unless_stream_ordered():
    """Without a shared CUDA stream the phases are asynchronous, so the driver must synchronise before NCCL reads a buffer and
    after it wrote it; with `stream_ordered` (library and torch on the same stream) no synchronisation is issued between the
    phases at all."""
    for stream_ordered, per_collective in ((False, 2), (True, 0)):
        eng, syncs = _MockEngine(0), []
        sharding.run_sharded(eng, _SingleRankDist, as_tensor=lambda t: t, sync=lambda: syncs.append(len(eng.log)),
                             stream_ordered=stream_ordered)
        ncoll = sum(1 for v in eng.bufs.values() if v is not None)
        assert len(syncs) == per_collective * ncoll + 1            # + the final one
        assert syncs[-1] == len(eng.log)
