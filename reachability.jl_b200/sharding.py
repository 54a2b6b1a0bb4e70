"""Multi-GPU partitioning of the hot path: one process per GPU, `torch.distributed` plumbing.

* BFFPSV18: the recurrence `P_k[R_g, :] = P_{k-1}[R_g, :] Φ` only needs the rank's own rows and the replicated Φ, so the
  rows of interest are sharded in contiguous groups of whole partition blocks (a 2-D block is never split), balanced
  by row count.  There is NO collective inside the time loop; the boxes are all-gathered once at the end.
* GLGM06 ensembles (config C5): independent members, sharded by contiguous batch ranges, no collective at all.
* A single GLGM06 system shards its generator COLUMNS (`run_sharded`): every rank maps, scores and gathers only the
  generators it owns; the Girard selection is key logic and runs replicated on the all-reduced scores, so all ranks take
  identical decisions.  Because the steps are processed in batches, the exchange is three small sum all-reduces per
  BATCH of steps (scores + row-support masks, W box partials, R box partials), not per reduction.  At n = 270 this does
  not beat one GPU (SURVEY.md section 8e says so); it exists for large n * p.

The compute callable is injected so that the host logic is testable on CPU with the `gloo` backend
(tests/test_sharding_cpu.py); the product path passes `Context.bffpsv18_boxes` / `Context.ensemble`.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import numpy as np


def shard_blocks(block_sizes: Sequence[int], rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) range of blocks owned by `rank`: contiguous, whole blocks, balanced by ROW count (greedy prefix split)."""
    sizes = np.asarray(block_sizes, dtype=np.int64)
    total = int(sizes.sum())
    ends = np.cumsum(sizes)
    # block b goes to the rank whose row interval [total*r/world, total*(r+1)/world) contains the block's midpoint
    mids = ends - sizes / 2.0
    owner = np.minimum((mids * world / max(total, 1)).astype(np.int64), world - 1)
    mine = np.nonzero(owner == rank)[0]
    if len(mine) == 0:
        lo = int(np.searchsorted(owner, rank))
        return lo, lo
    return int(mine[0]), int(mine[-1]) + 1


def shard_rows(partition: Sequence[Sequence[int]], blocks: Sequence[int], rank: int, world: int) -> np.ndarray:
    """0-based coordinates (in `blocks` order) of the rows this rank computes.  `partition`/`blocks` are 1-based as
    in the reference (compute_dimensions.jl:1-15)."""
    sizes = [len(partition[b - 1]) for b in blocks]
    lo, hi = shard_blocks(sizes, rank, world)
    return np.asarray([d - 1 for b in blocks[lo:hi] for d in partition[b - 1]], dtype=np.int32)


def shard_batch(batch: int, rank: int, world: int) -> slice:
    return slice((batch * rank) // world, (batch * (rank + 1)) // world)


def allgather_boxes(dist, local: "torch.Tensor", counts: List[int]):
    """All-gather row-sharded results: `local` is (2, rows_local * N) [centres; radii] on the rank's device; `counts`
    holds every rank's row count.  Shards are padded to the largest one so that a plain equal-size all-gather (one
    NCCL ring/NVLS collective over NVLink on GPUs, gloo on CPU) does the job.  Returns (per-rank tensors, N)."""
    import torch
    mine = counts[dist.get_rank()]
    t = torch.tensor([local.shape[1] // mine if mine else 0], dtype=torch.int64, device=local.device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)         # a rank may own no rows: learn N from the others
    N = int(t.item())
    width = max(counts) * N
    padded = torch.zeros((2, width), dtype=local.dtype, device=local.device)
    padded[:, :local.shape[1]] = local
    out = [torch.empty((2, width), dtype=local.dtype, device=local.device) for _ in counts]
    dist.all_gather(out, padded)
    return [o[:, :c * N] for o, c in zip(out, counts)], N


def bffpsv18_sharded(compute: Callable[[np.ndarray], Tuple[np.ndarray, np.ndarray]], partition, blocks, dist, device="cpu"):
    """Run `compute(rows) -> (centers, radii)` ((rows, N) arrays, column k-1 = step k) on this rank's rows and
    assemble the full (all rows of `blocks`, N) result on every rank."""
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    counts = [len(shard_rows(partition, blocks, r, world)) for r in range(world)]
    rows = shard_rows(partition, blocks, rank, world)
    if len(rows):
        C, R = compute(rows)
        N = C.shape[1]
        local = torch.from_numpy(np.stack([np.asfortranarray(C).ravel(order="F"), np.asfortranarray(R).ravel(order="F")]))
    else:
        local = torch.empty((2, 0), dtype=torch.float64)
    local = local.to(device)
    parts, N = allgather_boxes(dist, local, counts)
    Cs = [p[0].cpu().numpy().reshape((c, N), order="F") for p, c in zip(parts, counts) if c]
    Rs = [p[1].cpu().numpy().reshape((c, N), order="F") for p, c in zip(parts, counts) if c]
    return np.vstack(Cs), np.vstack(Rs)


# which buffers are sum-all-reduced after which phase of a batch (include/reach.h, reach_glgm06_shard_phase)
SHARD_COLLECTIVES = {0: (0, 1), 1: (2,), 2: (3,), 3: ()}


def _device_tensor(ptr, count, typestr, device):
    import torch

    class _Arr:
        pass
    a = _Arr()
    a.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(a, device=device)


def run_sharded(engine, dist, as_tensor=None, sync=None, stream_ordered=None, gemm_stream=None):
    """Drive a column-sharded GLGM06 flowpipe: `engine` exposes nbatches(), phase(batch, phase) and
    buffer(batch, which) (backend.ShardedFlowpipe, or a mock on CPU); `as_tensor` turns what buffer() returns into a
    torch tensor aliasing the engine's memory (default: a zero-copy view of the device buffer).

    `stream_ordered`: the engine's context issues its kernels on torch's CURRENT CUDA stream (``Context(dev,
    stream=torch.cuda.current_stream().cuda_stream)``).  The NCCL all-reduces are then ordered against the phases on the
    device (torch makes the collective wait for the current stream and the current stream wait for the collective), and
    the host enqueues the whole flowpipe without a single synchronisation.  Default: detected from the engine's context;
    otherwise the stream is synchronised around every collective.

    `gemm_stream` (a second ``torch.cuda.Stream``; needs `stream_ordered`): software pipeline across batches.  Phase 0 of batch
    b + 1 -- the map of the rank's own generator columns (the DMMA GEMM for a dense Phi) and their scores -- is issued on
    `gemm_stream` BEFORE phases 1..3 of batch b are issued on the current stream, so the dominant GEMM overlaps the selection
    chain, the box sums and their all-reduces of the previous batch (reach_glgm06_shard_phase_on; the library orders the reuse
    of its batch-local arrays).  The all-reduce of the scores of batch b + 1 is issued AFTER the collectives of batch b, so one
    NCCL communicator keeps its collectives in an order that never makes batch b wait for the GEMM of batch b + 1."""
    import torch
    if as_tensor is None:
        dev = torch.device("cuda", torch.cuda.current_device())

        def as_tensor(buf):
            return _device_tensor(buf[0], buf[1], buf[2], dev)
    if stream_ordered is None:
        ctx = getattr(engine, "ctx", None)
        stream_ordered = bool(torch.cuda.is_available() and ctx is not None and getattr(ctx, "stream", None) is not None
                              and ctx.stream == torch.cuda.current_stream().cuda_stream)
    if sync is None:
        def sync():
            if torch.cuda.is_available():
                torch.cuda.synchronize()
    if gemm_stream is not None:
        if not stream_ordered:
            raise ValueError("run_sharded: gemm_stream needs a context created on torch's current stream (stream_ordered)")
        main = torch.cuda.current_stream()
        nb = engine.nbatches()

        def reduce_buffers(b, phase, async_op=False):
            works = []
            for which in SHARD_COLLECTIVES[phase]:
                buf = engine.buffer(b, which)
                if buf is not None:
                    w = dist.all_reduce(as_tensor(buf), op=dist.ReduceOp.SUM, async_op=async_op)
                    if async_op and w is not None:
                        works.append(w)
            return works

        gemm_stream.wait_stream(main)
        engine.phase(0, 0, stream=gemm_stream.cuda_stream)
        for b in range(nb):
            with torch.cuda.stream(gemm_stream):
                # scores / masks of batch b (its phase 0 is in flight on gemm_stream).  ASYNC: the collective waits for
                # gemm_stream, but gemm_stream does not wait for the collective -- NCCL runs its collectives in issue order, so a
                # blocking all-reduce here would hold the GEMM of batch b + 1 back until phases 1..3 of batch b - 1 are through
                works = reduce_buffers(b, 0, async_op=True)
                ready = torch.cuda.Event()
                ready.record(gemm_stream)                               # the kernels of phase 0 of batch b
                if b + 1 < nb:
                    engine.phase(b + 1, 0, stream=gemm_stream.cuda_stream)     # overlaps everything below
            main.wait_event(ready)
            for w in works:
                w.wait()                                                # the MAIN stream waits for the all-reduced scores
            for phase in (1, 2, 3):
                engine.phase(b, phase, stream=main.cuda_stream)
                reduce_buffers(b, phase)
        main.wait_stream(gemm_stream)
        sync()
        return engine
    for b in range(engine.nbatches()):
        for phase in range(4):
            engine.phase(b, phase)
            for which in SHARD_COLLECTIVES[phase]:
                buf = engine.buffer(b, which)
                if buf is not None:
                    if not stream_ordered:
                        sync()                      # the phase's kernels must be done before NCCL reads the buffer
                    dist.all_reduce(as_tensor(buf), op=dist.ReduceOp.SUM)
                    if not stream_ordered:
                        sync()
    sync()
    return engine


def gather_sharded_step(engine, k, p, dist, device):
    """Reach set k of a column-sharded flowpipe on every rank: the ranks' slabs have disjoint support, so the sum is the
    generator matrix; centres are replicated."""
    import torch
    c, G = engine.step(k, p)
    t = torch.from_numpy(np.ascontiguousarray(G)).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return c, t.cpu().numpy()
