"""GPU (one device): the column-sharded GLGM06 phase machine (reach_glgm06_create_sharded / _shard_phase / _shard_buffer).
Two (or three) ranks are emulated on ONE GPU as separate sharded flowpipes whose all-reduce buffers are summed in
place between the phases -- exactly what `sharding.run_sharded` does with NCCL on a multi-GPU box
(tools/glgm06_sharded_check.py is the real multi-process run; profiles/r01/glgm06_sharded_2gpu_check.log).
Results must equal the oracle: every rank takes the same selections from the all-reduced scores, the ranks' slabs have
disjoint support and sum to the reach set."""
import numpy as np
import pytest
import torch

from oracle import reach_oracle as orc
from reachability_jl_b200 import sharding
from tests.util import assert_close

pytestmark = pytest.mark.gpu


class _EmulatedRanks:
    """Drives several sharded engines of one process like sharding.run_sharded drives one per process."""

    def __init__(self, engines):
        self.engines = engines
        self.dev = torch.device("cuda", 0)

    def run(self):
        for b in range(self.engines[0].nbatches()):
            for phase in range(4):
                for e in self.engines:
                    e.phase(b, phase)
                self.engines[0].ctx.sync()       # the phases are asynchronous; the emulated collective runs on torch's stream
                for which in sharding.SHARD_COLLECTIVES[phase]:
                    bufs = [e.buffer(b, which) for e in self.engines]
                    if bufs[0] is None:
                        assert all(x is None for x in bufs)
                        continue
                    ts = [sharding._device_tensor(p, c, d, self.dev) for (p, c, d) in bufs]
                    total = torch.stack(ts).sum(dim=0)
                    for t in ts:
                        t.copy_(total)
                torch.cuda.synchronize()

    def _reduce(self, b, phase):
        for which in sharding.SHARD_COLLECTIVES[phase]:
            bufs = [e.buffer(b, which) for e in self.engines]
            if bufs[0] is None:
                assert all(x is None for x in bufs)
                continue
            ts = [sharding._device_tensor(p, c, d, self.dev) for (p, c, d) in bufs]
            total = torch.stack(ts).sum(dim=0)
            for t in ts:
                t.copy_(total)

    def run_pipelined(self):
        """The order of sharding.run_sharded(..., gemm_stream=...): phase 0 of batch b + 1 on a second stream, issued before
        phases 1..3 of batch b; the emulated collectives run on the stream of their phase, nothing synchronises the host."""
        sg, sm = torch.cuda.Stream(), torch.cuda.Stream()
        nb = self.engines[0].nbatches()
        torch.cuda.synchronize()
        for e in self.engines:
            e.phase(0, 0, stream=sg.cuda_stream)
        for b in range(nb):
            with torch.cuda.stream(sg):
                self._reduce(b, 0)
                ready = torch.cuda.Event()
                ready.record(sg)
            if b + 1 < nb:
                for e in self.engines:
                    e.phase(b + 1, 0, stream=sg.cuda_stream)
            sm.wait_event(ready)
            with torch.cuda.stream(sm):
                for phase in (1, 2, 3):
                    for e in self.engines:
                        e.phase(b, phase, stream=sm.cuda_stream)
                    self._reduce(b, phase)
        torch.cuda.synchronize()

    def step(self, k):
        p = self.engines[0].ngens()
        parts = [e.step(k, int(p[k - 1])) for e in self.engines]
        for (c, _) in parts[1:]:
            assert np.array_equal(c, parts[0][0])            # centres are replicated
        G = sum(g for (_, g) in parts)
        support = sum((np.abs(g).sum(axis=0) > 0).astype(int) for (_, g) in parts)
        assert support.max() <= 1                            # disjoint column support
        return parts[0][0], G


def _problem(seed, n, m):
    g = np.random.default_rng(seed)
    A = g.standard_normal((n, n)) / np.sqrt(n) - 1.2 * np.eye(n)
    X0 = orc.Box(g.standard_normal(n), np.abs(g.standard_normal(n)) * 0.1)
    B = g.standard_normal((n, m))
    U = orc.Box(g.standard_normal(m), np.abs(g.standard_normal(m)) * 0.3)
    return A, X0, B, U


@pytest.mark.parametrize("nranks", [1, 2, 3])
@pytest.mark.parametrize("n,m,N,r,rzg", [(6, 1, 40, 2, True), (13, 3, 33, 4, True), (20, 20, 20, 5, False), (33, 2, 70, 6, True),
                                         (9, 2, 60, 1, True)])
def test_column_sharded_flowpipe_equals_oracle(ctx, nranks, n, m, N, r, rzg):
    A, X0, B, U = _problem(700 + n, n, m)
    D = orc.discretize_zonotope(A, X0, 0.01, B, U, rzg=rzg)
    Om = orc.reduce_order(D.Omega0, r, rzg)
    R = orc.glgm06_reach_inhomog(Om, D.V, D.Phi, N, r, rzg)
    engines = [ctx.glgm06_sharded(D.Phi, Om.c, Om.G, N, r, D.V.c, D.V.G, rk, nranks, flags=0 if rzg else 1) for rk in range(nranks)]
    emu = _EmulatedRanks(engines)
    emu.run()
    emu.run()                                # the phase machine restarts cleanly (idempotent)
    p = engines[0].ngens()
    for e in engines[1:]:
        assert np.array_equal(e.ngens(), p)  # replicated selection: identical generator counts on every rank
    for k in range(1, N + 1):
        assert p[k - 1] == R[k - 1].ngens
        c, G = emu.step(k)
        assert_close(c, R[k - 1].c, f"centre {k}")
        assert_close(G, R[k - 1].G, f"generators {k}")
    for e in engines:
        e.close()


def test_sharded_engine_refuses_plain_run(ctx):
    import reachability_jl_b200 as rb
    A, X0, B, U = _problem(1, 5, 1)
    D = orc.discretize_zonotope(A, X0, 0.01, B, U)
    e = ctx.glgm06_sharded(D.Phi, D.Omega0.c, D.Omega0.G, 5, 3, D.V.c, D.V.G, 0, 2)
    with pytest.raises(rb.ReachError):
        e.run()
    with pytest.raises(rb.ReachError):
        e.phase(2, 0)                        # batches must be processed in order
    e.close()


def test_column_sharded_flowpipe_with_structurally_sparse_phi(ctx):
    """Decoupled 2 x 2 modes: every rank maps its generators with the ELL sparse product instead of the GEMM."""
    g = np.random.default_rng(21)
    n, m, N, r, nranks = 36, 2, 50, 3, 2
    A = np.zeros((n, n))
    Phi = np.zeros((n, n))
    for i in range(n // 2):
        w = g.uniform(1.0, 30.0)
        blk = np.array([[-0.1, w], [-w, -0.1]])
        A[2 * i:2 * i + 2, 2 * i:2 * i + 2] = blk
        Phi[2 * i:2 * i + 2, 2 * i:2 * i + 2] = orc.exp_Adelta(blk, 0.01)
    X0 = orc.Box(g.standard_normal(n), np.abs(g.standard_normal(n)) * 0.1)
    B = g.standard_normal((n, m))
    U = orc.Box(g.standard_normal(m), np.abs(g.standard_normal(m)) * 0.3)
    D = orc.discretize_zonotope(A, X0, 0.01, B, U, Phi=Phi)
    Om = orc.reduce_order(D.Omega0, r)
    R = orc.glgm06_reach_inhomog(Om, D.V, D.Phi, N, r)
    engines = [ctx.glgm06_sharded(D.Phi, Om.c, Om.G, N, r, D.V.c, D.V.G, rk, nranks) for rk in range(nranks)]
    assert all(e.map_info()["nnz_per_row"] == 2 for e in engines)
    emu = _EmulatedRanks(engines)
    emu.run()
    p = engines[0].ngens()
    for k in range(1, N + 1):
        assert p[k - 1] == R[k - 1].ngens
        c, G = emu.step(k)
        assert_close(c, R[k - 1].c, f"centre {k}")
        assert_close(G, R[k - 1].G, f"generators {k}")
    for e in engines:
        e.close()


@pytest.mark.parametrize("nranks", [1, 2, 3])
@pytest.mark.parametrize("n,m,N,r", [(33, 2, 330, 3), (64, 5, 200, 4)])
def test_pipelined_phase_order_equals_serial_phase_order(ctx, nranks, n, m, N, r):
    """reach_glgm06_shard_phase_on: phase 0 of batch b + 1 issued on a second stream before phases 1..3 of batch b (>= 4 batches,
    so every batch-local array set is reused) gives the bits of the in-order run, and repeats them when run again."""
    A, X0, B, U = _problem(500 + n, n, m)
    D = orc.discretize_zonotope(A, X0, 0.01, B, U)
    Om = orc.reduce_order(D.Omega0, r)
    mk = lambda: [ctx.glgm06_sharded(D.Phi, Om.c, Om.G, N, r, D.V.c, D.V.G, rk, nranks) for rk in range(nranks)]
    ser, pip = _EmulatedRanks(mk()), _EmulatedRanks(mk())
    ser.run()
    pip.run_pipelined()
    pip.run_pipelined()                  # a second run starts from scratch
    p = ser.engines[0].ngens()
    assert np.array_equal(p, pip.engines[0].ngens())
    for k in sorted(set([1, 2, 3, 64, 65, 129, 200, N - 1, N]) & set(range(1, N + 1))):
        cs, Gs = ser.step(k)
        cp, Gp = pip.step(k)
        assert np.array_equal(cs, cp) and np.array_equal(Gs, Gp), k
    Ro = orc.glgm06_reach_inhomog(Om, D.V, D.Phi, N, r)
    c, G = pip.step(N)
    assert G.shape == Ro[N - 1].G.shape
    assert_close(c, Ro[N - 1].c, "pipelined sharded centre")
    assert_close(G, Ro[N - 1].G, "pipelined sharded generators")
    for e in ser.engines + pip.engines:
        e.close()


def test_run_sharded_with_a_gemm_stream_single_rank(ctx_on_torch_stream):
    """sharding.run_sharded(..., gemm_stream=...) itself (one rank, no-op collective) against the oracle."""
    ctx, stream = ctx_on_torch_stream

    class _Solo:
        class ReduceOp:
            SUM = None

        @staticmethod
        def all_reduce(t, op=None, async_op=False):
            return None

    A, X0, B, U = _problem(77, 40, 3)
    D = orc.discretize_zonotope(A, X0, 0.01, B, U)
    Om = orc.reduce_order(D.Omega0, 3)
    N = 280
    with torch.cuda.stream(stream):
        e = ctx.glgm06_sharded(D.Phi, Om.c, Om.G, N, 3, D.V.c, D.V.G, 0, 1)
        sharding.run_sharded(e, _Solo, gemm_stream=torch.cuda.Stream())
        p = e.ngens()
        Ro = orc.glgm06_reach_inhomog(Om, D.V, D.Phi, N, 3)
        for k in (1, 2, 100, 257, N):
            c, G = e.step(k, int(p[k - 1]))
            assert G.shape == Ro[k - 1].G.shape
            assert_close(c, Ro[k - 1].c, f"centre {k}")
            assert_close(G, Ro[k - 1].G, f"generators {k}")
        e.close()
