"""GPU parity at FULL LENGTH: every reach set and every order-reduction selection of long GLGM06 flowpipes
(reach_inhomog!, GLGM06/reach.jl:30-62, through the C ABI) against the CPU oracle.

The pipelined device path processes the steps in batches (up to 64 steps) and rotates THREE sets of batch-local
arrays, so a set is first reused at the fourth batch.  The short parity tests (tests/test_glgm06_gpu.py, N <= 70)
never get there; these do: N >= 300 with batches of 8 and of 64 steps, BASELINE.json's C2 at its full N = 2000 on the
sparse-Phi (ELL) map and on the dense DMMA GEMM, and the column-sharded engine with emulated ranks.

What is compared, for EVERY step k = 1..N: the generator count, the centre and every generator entry within
1e-10 relative + 1e-12 absolute, and for both reduce_order calls of the step (R_k :49, W :55) the number of boxed
generators m, the full sortperm and the scores.  The oracle is streamed (oracle.glgm06_iter_inhomog) so that the
11.7 GB flowpipe of C2 is never held twice on the host.
"""
import numpy as np
import pytest

from oracle import reach_oracle as orc
from tests.util import tol_units

pytestmark = pytest.mark.gpu

KEEP, RECORD = 1, 2


def _check_selection(got, sel, k, which, label):
    if sel is None:
        assert got["p_in"] == 0 and got["m"] == 0, f"{label} step {k} list {which}: unexpected reduction"
        return
    h, ind, m = sel
    assert got["p_in"] == len(h) and got["m"] == m, (label, k, which, got["p_in"], len(h), got["m"], m)
    if not np.array_equal(got["ind"], ind):
        i0 = int(np.nonzero(got["ind"] != ind)[0][0])
        raise AssertionError(f"{label} step {k} list {which}: sortperm differs at rank {i0} (got {got['ind'][i0]}, oracle "
                             f"{ind[i0]}; h = {h[got['ind'][i0]]!r} vs {h[ind[i0]]!r})")
    u = tol_units(got["h"], h)
    assert u <= 1.0, f"{label} step {k} list {which}: scores off by {u:.3g} tolerance units"


def _stream_compare(readers, Om, V, Phi, N, r, rzg=True):
    """readers: {label: (ngens array, step(k) -> (c, G), selection(k, which) or None)}.  Returns the worst entrywise
    deviation in tolerance units per label."""
    worst = {label: 0.0 for label in readers}
    for k, Zo, sel in orc.glgm06_iter_inhomog(Om, V, Phi, N, r, rzg):
        for label, (p, step, selection) in readers.items():
            assert p[k - 1] == Zo.ngens, f"{label} step {k}: {p[k - 1]} generators, oracle {Zo.ngens}"
            c, G = step(k)
            uc, uG = tol_units(c, Zo.c), tol_units(G, Zo.G)
            assert uc <= 1.0, f"{label} step {k} centre: {uc:.3g} tolerance units"
            if uG > 1.0:
                j = int(np.argmax(np.max(np.abs(G - Zo.G) / (1e-10 * np.abs(Zo.G) + 1e-12), axis=0)))
                raise AssertionError(f"{label} step {k} generators: {uG:.3g} tolerance units (worst column {j} of {Zo.ngens})")
            worst[label] = max(worst[label], uc, uG)
            if selection is not None and sel is not None:
                for which in (0, 1):
                    _check_selection(selection(k, which), sel[which], k, which, label)
    return worst


def _fp_reader(fp):
    p = fp.ngens()
    return p, (lambda k: fp.step(k, int(p[k - 1]))), fp.selection


def _dense_problem(seed, n, m, delta=0.01):
    g = np.random.default_rng(seed)
    A = g.standard_normal((n, n)) / np.sqrt(n) - 1.2 * np.eye(n)
    X0 = orc.Box(g.standard_normal(n), np.abs(g.standard_normal(n)) * 0.1)
    B = g.standard_normal((n, m))
    U = orc.Box(g.standard_normal(m), np.abs(g.standard_normal(m)) * 0.3)
    D = orc.discretize_zonotope(A, X0, delta, B, U)
    return D


@pytest.mark.parametrize("block", [8, 64])
@pytest.mark.parametrize("n,m,N,r", [(33, 2, 420, 3), (96, 5, 330, 4), (270, 3, 300, 3)])
def test_long_dense_flowpipes_every_step_and_selection(ctx, n, m, N, r, block, monkeypatch):
    """Random DENSE systems (the DMMA GEMM map), N >= 300: with batches of 8 steps the three rotating array sets are
    reused ~13 times over, with batches of 64 the fourth batch (step >= 193) is the first reuse."""
    monkeypatch.setenv("REACH_GLGM06_BLOCK", str(block))
    D = _dense_problem(900 + n, n, m)
    Om = orc.reduce_order(D.Omega0, r)
    fp = ctx.glgm06(D.Phi, Om.c, Om.G, N, r, D.V.c, D.V.G, RECORD).run()
    assert fp.map_info()["sparse_levels"] == 0
    info = fp.chain_info()
    assert info["batches"] >= 4 and info["parallel_batches"] >= info["batches"] - 8, info      # all but the growth phase
    worst = _stream_compare({f"dense n={n} block={block}": _fp_reader(fp)}, Om, D.V, D.Phi, N, r)
    assert max(worst.values()) <= 1.0
    fp.close()


@pytest.mark.parametrize("n,m,N,r", [(33, 2, 420, 3), (96, 5, 330, 4)])
def test_parallel_selection_chain_is_bit_identical_to_the_sequential_chain(ctx, n, m, N, r, monkeypatch):
    """The parallel chain (every generator's fate over a batch evaluated independently from counts) and the sequential
    merge chain (REACH_CHAIN_SEQUENTIAL=1) must take the same selections in the same order: every generator, box vector
    and recorded sortperm agrees BIT FOR BIT."""
    D = _dense_problem(900 + n, n, m)
    Om = orc.reduce_order(D.Omega0, r)
    fpar = ctx.glgm06(D.Phi, Om.c, Om.G, N, r, D.V.c, D.V.G, RECORD).run()
    assert fpar.chain_info()["parallel_batches"] >= 3
    monkeypatch.setenv("REACH_CHAIN_SEQUENTIAL", "1")
    fseq = ctx.glgm06(D.Phi, Om.c, Om.G, N, r, D.V.c, D.V.G, RECORD).run()
    monkeypatch.delenv("REACH_CHAIN_SEQUENTIAL")
    assert fseq.chain_info()["parallel_batches"] == 0
    p = fpar.ngens()
    assert np.array_equal(p, fseq.ngens())
    for k in range(1, N + 1):
        ca, Ga = fpar.step(k, int(p[k - 1]))
        cb, Gb = fseq.step(k, int(p[k - 1]))
        assert np.array_equal(ca, cb) and np.array_equal(Ga, Gb), f"step {k}"
        if k >= 2:
            for which in (0, 1):
                sa, sb = fpar.selection(k, which), fseq.selection(k, which)
                assert sa["p_in"] == sb["p_in"] and sa["m"] == sb["m"] and np.array_equal(sa["ind"], sb["ind"]), (k, which)
                assert np.array_equal(sa["h"], sb["h"])
    fpar.close(); fseq.close()


@pytest.mark.parametrize("rzg", [True, False])
def test_long_structured_flowpipe_with_zero_generator_removal(ctx, rzg, monkeypatch):
    """Decoupled modes, inputs entering two rows, part of X0 flat: box vectors keep exact zeros, so generator counts are
    data dependent for hundreds of steps (zero-generator removal) -- batches of 16 steps, N = 400."""
    monkeypatch.setenv("REACH_GLGM06_BLOCK", "16")
    g = np.random.default_rng(77)
    n, m, N, r = 40, 2, 400, 3
    A = np.zeros((n, n))
    Phi = np.zeros((n, n))
    for i in range(n // 2):
        w = g.uniform(1.0, 30.0)
        blk = np.array([[-0.1, w], [-w, -0.1]])
        A[2 * i:2 * i + 2, 2 * i:2 * i + 2] = blk
        Phi[2 * i:2 * i + 2, 2 * i:2 * i + 2] = orc.exp_Adelta(blk, 0.01)
    X0 = orc.Box(g.standard_normal(n), np.abs(g.standard_normal(n)) * 0.1)
    X0.r[n // 2:] = 0.0
    B = np.zeros((n, m)); B[1, 0] = 1.0; B[n // 2 | 1, 1] = -0.5
    U = orc.Box(g.standard_normal(m), np.abs(g.standard_normal(m)) * 0.3)
    D = orc.discretize_zonotope(A, X0, 0.01, B, U, rzg=rzg, Phi=Phi)
    Om = orc.reduce_order(D.Omega0, r, rzg)
    fp = ctx.glgm06(D.Phi, Om.c, Om.G, N, r, D.V.c, D.V.G, RECORD | (0 if rzg else KEEP)).run()
    worst = _stream_compare({f"structured rzg={rzg}": _fp_reader(fp)}, Om, D.V, D.Phi, N, r, rzg)
    assert max(worst.values()) <= 1.0
    fp.close()


def _c2():
    from reachability_jl_b200 import workloads as wl
    w = wl.iss()
    Phi = orc.exp_Adelta(w.A, w.delta)
    P2 = orc.phi12_series(np.abs(w.A), w.delta)[2]
    D = orc.discretize_zonotope(w.A, orc.Box(w.c, w.r), w.delta, w.B, orc.Box(w.cU, w.rU), Phi=Phi, Phi2abs=P2)
    Om = orc.reduce_order(D.Omega0, w.max_order)
    return w, D, Om, orc.n_steps_glgm06(w.T, w.delta)


def test_config_c2_full_length_every_step_sparse_and_dense_map(ctx, monkeypatch):
    """BASELINE.json configs[1] at full size -- n = 270, max_order 10, N = 2000 (32 batches, 2700 generators per reach
    set) -- on BOTH generator maps of the final pipelined code: the ELL sparse product fused with the scores (what
    bench.py times as the C2 value) and the dense DMMA GEMM (REACH_GLGM06_DENSE=1).  All 2000 reach sets entrywise, all
    3998 sortperms."""
    w, D, Om, N = _c2()
    n, r = w.n, w.max_order
    assert N == 2000 and Om.ngens == 1084 and D.V.ngens == 273
    fs = ctx.glgm06(D.Phi, Om.c, Om.G, N, r, D.V.c, D.V.G, RECORD).run()
    assert fs.map_info()["sparse_levels"] >= 1
    assert fs.chain_info()["parallel_batches"] >= 25        # 38 batches; all but the growth phase take the parallel selection chain
    monkeypatch.setenv("REACH_GLGM06_DENSE", "1")
    fd = ctx.glgm06(D.Phi, Om.c, Om.G, N, r, D.V.c, D.V.G, RECORD).run()
    monkeypatch.delenv("REACH_GLGM06_DENSE")
    assert fd.map_info()["sparse_levels"] == 0
    worst = _stream_compare({"C2 sparse map": _fp_reader(fs), "C2 dense GEMM": _fp_reader(fd)}, Om, D.V, D.Phi, N, r)
    print("C2 full length, worst tolerance units:", worst)
    assert max(worst.values()) <= 1.0
    # a second run of the same handle reproduces the last reach set bit for bit (fixed summation orders, no float atomics)
    c_a, G_a = fs.step(N)
    fs.run()
    c_b, G_b = fs.step(N)
    assert np.array_equal(c_a, c_b) and np.array_equal(G_a, G_b)
    fs.close(); fd.close()


@pytest.mark.parametrize("n,m,N,r,nranks", [(33, 2, 320, 3, 2), (96, 5, 300, 4, 3)])
def test_long_column_sharded_flowpipes_every_step(ctx, n, m, N, r, nranks):
    """The column-sharded phase machine (reach_glgm06_create_sharded) with emulated ranks over >= 4 batches: the summed
    slabs equal the oracle at every step and the replicated selections equal the oracle's sortperms."""
    from tests.test_glgm06_sharded_gpu import _EmulatedRanks
    D = _dense_problem(950 + n, n, m)
    Om = orc.reduce_order(D.Omega0, r)
    engines = [ctx.glgm06_sharded(D.Phi, Om.c, Om.G, N, r, D.V.c, D.V.G, rk, nranks, flags=RECORD) for rk in range(nranks)]
    emu = _EmulatedRanks(engines)
    emu.run()
    p = engines[0].ngens()
    for e in engines[1:]:
        assert np.array_equal(e.ngens(), p)
    worst = _stream_compare({f"sharded x{nranks} n={n}": (p, emu.step, engines[0].selection)}, Om, D.V, D.Phi, N, r)
    assert max(worst.values()) <= 1.0
    for e in engines:
        e.close()
