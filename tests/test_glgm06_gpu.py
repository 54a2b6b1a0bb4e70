"""GPU parity: GLGM06 device path (reach_homog! / reach_inhomog!, GLGM06/reach.jl:5-62, through the C ABI) vs the
CPU oracle.

Tolerance (north_star): 1e-10 relative + 1e-12 absolute on every centre and generator entry, and IDENTICAL
order-reduction selections (sortperm and number of boxed generators of every reduce_order call).
Shapes mirror /root/reference/test/Reachability/solve_continuous.jl:186-196 (4x4 LCS / CLCCS with GLGM06 and
`:max_order`) plus random dense systems with ragged sizes and BASELINE.json's C2-shaped system at reduced size.
"""
import numpy as np
import pytest

from oracle import reach_oracle as orc
from tests.util import assert_close, tol_units

pytestmark = pytest.mark.gpu

KEEP, RECORD = 1, 2


def _problem(seed, n, m, delta=0.01, dense=True):
    g = np.random.default_rng(seed)
    if dense:
        A = g.standard_normal((n, n)) / np.sqrt(n) - 1.2 * np.eye(n)
    else:   # block-diagonal 2x2 rotations (ISS-shaped): generators keep structural zeros
        A = np.zeros((n, n))
        for i in range(n // 2):
            w = g.uniform(1.0, 30.0)
            A[2 * i:2 * i + 2, 2 * i:2 * i + 2] = [[-0.1, w], [-w, -0.1]]
        if n % 2:
            A[-1, -1] = -1.0
    X0 = orc.Box(g.standard_normal(n), np.abs(g.standard_normal(n)) * 0.1)
    B = g.standard_normal((n, m)) if m else None
    U = orc.Box(g.standard_normal(m), np.abs(g.standard_normal(m)) * 0.3) if m else None
    return A, X0, B, U, delta


def _oracle(A, X0, B, U, delta, N, r, rzg):
    D = orc.discretize_zonotope(A, X0, delta, B, U, rzg=rzg)
    Om = orc.reduce_order(D.Omega0, r, rzg)
    if D.V is None:
        return D, Om, orc.glgm06_reach_homog(Om, D.Phi, N, rzg), [None] * N
    R, sels = orc.glgm06_reach_inhomog(Om, D.V, D.Phi, N, r, rzg, return_selections=True)
    return D, Om, R, sels


def _device(ctx, D, Om, N, r, rzg, record=True):
    flags = (0 if rzg else KEEP) | (RECORD if record else 0)
    V = D.V
    fp = ctx.glgm06(D.Phi, Om.c, Om.G, N, r, None if V is None else V.c, None if V is None else V.G, flags)
    fp.run()
    return fp


def _compare(fp, R, sels, N, check_sel=True):
    p = fp.ngens()
    for k in range(1, N + 1):
        Zo = R[k - 1]
        assert p[k - 1] == Zo.ngens, f"step {k}: {p[k - 1]} generators, oracle {Zo.ngens}"
        c, G = fp.step(k)
        assert_close(c, Zo.c, f"step {k} centre")
        if check_sel and sels[k - 1] is not None:
            for which, sel in enumerate(sels[k - 1]):
                got = fp.selection(k, which)
                if sel is None:
                    assert got["p_in"] == 0 and got["m"] == 0, f"step {k} list {which}: unexpected reduction"
                    continue
                h, ind, m = sel
                assert got["p_in"] == len(h) and got["m"] == m, (k, which, got["p_in"], len(h), got["m"], m)
                if not np.array_equal(got["ind"], ind):
                    bad = np.nonzero(got["ind"] != ind)[0]
                    i0 = bad[0]
                    raise AssertionError(f"step {k} list {which}: sortperm differs at rank {i0} "
                                         f"(got {got['ind'][i0]}, oracle {ind[i0]}; h={h[got['ind'][i0]]!r} vs {h[ind[i0]]!r})")
                assert tol_units(got["h"], h) <= 1.0
        assert_close(G, Zo.G, f"step {k} generators")


@pytest.mark.parametrize("n,N", [(1, 4), (4, 20), (13, 33), (28, 65), (50, 40)])
def test_homogeneous(ctx, n, N):
    A, X0, _, _, delta = _problem(200 + n, n, 0)
    D, Om, R, sels = _oracle(A, X0, None, None, delta, N, 10, True)
    fp = _device(ctx, D, Om, N, 10, True, record=False)
    _compare(fp, R, sels, N, check_sel=False)
    lo, hi = fp.box(N)
    rad = np.abs(R[-1].G).sum(axis=1)
    assert_close(lo, R[-1].c - rad, "box lo")
    assert_close(hi, R[-1].c + rad, "box hi")
    fp.close()


@pytest.mark.parametrize("rzg", [True, False])
@pytest.mark.parametrize("n,m,N,r", [(4, 2, 30, 3), (6, 1, 40, 2), (13, 3, 25, 4), (20, 20, 20, 5), (33, 2, 30, 6),
                                     (64, 5, 20, 3), (9, 2, 60, 1)])
def test_inhomogeneous_dense(ctx, n, m, N, r, rzg):
    A, X0, B, U, delta = _problem(300 + n, n, m)
    D, Om, R, sels = _oracle(A, X0, B, U, delta, N, r, rzg)
    fp = _device(ctx, D, Om, N, r, rzg)
    _compare(fp, R, sels, N)
    fp.close()


@pytest.mark.parametrize("rzg", [True, False])
@pytest.mark.parametrize("n,m,N,r", [(8, 1, 40, 3), (20, 2, 30, 4), (54, 3, 30, 5)])
def test_inhomogeneous_structured_zeros(ctx, n, m, N, r, rzg):
    """Decoupled modes with the input entering few rows: box vectors d have exact zeros, so zero-generator removal
    (`Zonotope` constructor) changes the generator counts and the number of boxed generators."""
    A, X0, B, U, delta = _problem(400 + n, n, m, dense=False)
    B = np.zeros((n, m))
    B[1, 0] = 1.0
    if m > 1:
        B[n // 2 | 1, 1] = -0.5
    X0.r[n // 2:] = 0.0
    D, Om, R, sels = _oracle(A, X0, B, U, delta, N, r, rzg)
    fp = _device(ctx, D, Om, N, r, rzg)
    _compare(fp, R, sels, N)
    fp.close()


def test_n_equals_one_set(ctx):
    A, X0, B, U, delta = _problem(7, 5, 2)
    D, Om, R, sels = _oracle(A, X0, B, U, delta, 1, 3, True)
    fp = _device(ctx, D, Om, 1, 3, True)
    _compare(fp, R, sels, 1)
    fp.close()


def test_bulk_fetch_matches_stepwise(ctx):
    A, X0, B, U, delta = _problem(11, 16, 2)
    N = 25
    D, Om, R, sels = _oracle(A, X0, B, U, delta, N, 3, True)
    fp = _device(ctx, D, Om, N, 3, True)
    C, G, p = fp.fetch()
    for k in range(1, N + 1):
        c, Gk = fp.step(k)
        assert np.array_equal(C[:, k - 1], c)
        assert np.array_equal(G[:, :p[k - 1], k - 1], Gk)
    fp.close()


def test_streamed_run_fetch_matches_fetch(ctx):
    A, X0, B, U, delta = _problem(13, 20, 3)
    N = 70
    D, Om, R, sels = _oracle(A, X0, B, U, delta, N, 4, True)
    fp = _device(ctx, D, Om, N, 4, True, record=False)
    C, G, p = fp.fetch()
    C2 = np.full_like(C, np.nan)
    G2 = np.full_like(G, np.nan)
    fp.run_fetch(C2, G2)
    assert np.array_equal(C, C2)
    for k in range(N):
        assert np.array_equal(G[:, :p[k], k], G2[:, :p[k], k])
        assert np.all(G2[:, p[k]:, k] == 0.0)
    fp.close()


def test_c2_shaped_reduced(ctx):
    """BASELINE.json config C2 (ISS-shaped modal system, GLGM06, max_order 10) at n=54, N=60."""
    from reachability_jl_b200 import workloads as wl
    w = wl.iss(n_modes=27, T=0.6)
    N = orc.n_steps_glgm06(w.T, w.delta)
    D, Om, R, sels = _oracle(w.A, orc.Box(w.c, w.r), w.B, orc.Box(w.cU, w.rU), w.delta, N, w.max_order, True)
    fp = _device(ctx, D, Om, N, w.max_order, True)
    _compare(fp, R, sels, N)
    fp.close()


def test_run_is_idempotent(ctx):
    A, X0, B, U, delta = _problem(12, 10, 2)
    N = 30
    D, Om, R, sels = _oracle(A, X0, B, U, delta, N, 3, True)
    fp = _device(ctx, D, Om, N, 3, True)
    fp.run()
    fp.run()
    _compare(fp, R, sels, N)
    fp.close()


@pytest.mark.parametrize("rzg", [True, False])
@pytest.mark.parametrize("n,p,r", [(3, 10, 2), (5, 40, 3), (16, 100, 1), (16, 100, 5), (16, 100, 7), (33, 500, 4), (270, 3784, 10),
                                   (7, 3, 2)])
def test_reduce_order_standalone(ctx, n, p, r, rzg):
    """`reduce_order(Z, r)` (LazySets, call site GLGM06/post.jl:20): Girard selection, kept generators in
    ascending-score order, box generators last; unchanged when r*n >= p."""
    g = np.random.default_rng(n * 1000 + p)
    G = g.standard_normal((n, p)) * np.exp(g.uniform(-3, 3, p))[None, :]
    G[:, ::7] = 0.0
    G[n // 2, 3::7] = 0.0                # some structure: zero columns and zero entries
    if p > 10:
        G[:, 5] = 0.0
        G[0, 5] = 2.5                    # an axis-aligned generator (score exactly 0)
    c = g.standard_normal(n)
    Zin = orc.make_zono(c, G, rzg)
    Zo, sel = orc.reduce_order(Zin, r, rzg, return_selection=True)
    flags = 0 if rzg else KEEP
    Gd, ind = ctx.reduce_order(c, Zin.G, r, flags)
    assert Gd.shape == Zo.G.shape
    assert_close(Gd, Zo.G, "reduced generators")
    if sel is not None:
        assert np.array_equal(ind[:len(sel[1])], sel[1])


@pytest.mark.parametrize("n,p,batch,N", [(4, 5, 3, 10), (28, 85, 16, 40), (6, 0, 4, 7)])
def test_ensemble_homogeneous(ctx, n, p, batch, N):
    """Batched reach_homog! (GLGM06/reach.jl:15-19): every member must equal its own single flowpipe."""
    g = np.random.default_rng(n + p + batch)
    A = g.standard_normal((n, n)) / np.sqrt(n) - np.eye(n)
    Phi = orc.exp_Adelta(A, 0.05)
    c0s = g.standard_normal((n, batch))
    G0s = g.standard_normal((n, p, batch)) * 0.1
    ens = ctx.ensemble(Phi, c0s, G0s, N).run()
    for b in range(batch):
        R = orc.glgm06_reach_homog(orc.Zono(c0s[:, b], G0s[:, :, b]), Phi, N, rzg=False)
        for k in (1, 2, N // 2, N):
            c, G = ens.step(b, k)
            assert_close(c, R[k - 1].c, f"member {b} step {k} centre")
            assert_close(G, R[k - 1].G.reshape(n, p), f"member {b} step {k} generators")
    lo, hi = ens.boxes(N)
    for b in range(batch):
        R = orc.glgm06_reach_homog(orc.Zono(c0s[:, b], G0s[:, :, b]), Phi, N, rzg=False)
        rad = np.abs(R[-1].G).sum(axis=1) if p else np.zeros(n)
        assert_close(lo[:, b], R[-1].c - rad, "lo")
        assert_close(hi[:, b], R[-1].c + rad, "hi")
    # every member at every step in one pass (reach_ensemble_boxes_all): identical bits to the per-step query
    LO, HI = ens.boxes_all()
    for k in (1, 2, N):
        lk, hk = ens.boxes(k)
        assert np.array_equal(LO[:, :, k - 1], lk) and np.array_equal(HI[:, :, k - 1], hk)
    ens.close()


def test_long_w_list_uses_global_scratch(ctx):
    """r*n large enough that the selection chain's working set exceeds shared memory (global-scratch instantiation)."""
    A, X0, B, U, delta = _problem(21, 96, 40)
    N, r = 12, 75                                   # capW = 7208 entries > 200 KB of chain state
    D, Om, R, sels = _oracle(A, X0, B, U, delta, N, r, True)
    fp = _device(ctx, D, Om, N, r, True)
    _compare(fp, R, sels, N)
    fp.close()
    # and a reduction that actually triggers with a long list: many input generators, moderate order
    A, X0, B, U, delta = _problem(22, 64, 64)
    N, r = 40, 3
    D, Om, R, sels = _oracle(A, X0, B, U, delta, N, r, False)
    fp = _device(ctx, D, Om, N, r, False)
    _compare(fp, R, sels, N)
    fp.close()


@pytest.mark.parametrize("n,m,N,r,dense", [(1100, 2, 7, 2, True), (1500, 3, 6, 3, True), (1040, 2, 12, 2, False)])
def test_state_dimension_above_1024(ctx, n, m, N, r, dense):
    """GLGM06/reach.jl:30-62 has no limit on n.  The numerics kernels sweep the rows in windows of 1024, the sequential
    selection chain keeps per-row state for up to 4096 rows; the structured case (decoupled modes, inputs entering two rows)
    keeps zero-generator removal active with more than 32 mask words per generator."""
    A, X0, B, U, delta = _problem(800 + n, n, m, dense=dense)
    if not dense:
        B = np.zeros((n, m)); B[1, 0] = 1.0; B[n // 2 | 1, 1] = -0.5
        X0.r[n // 2:] = 0.0
    D, Om, R, sels = _oracle(A, X0, B, U, delta, N, r, True)
    fp = _device(ctx, D, Om, N, r, True)
    _compare(fp, R, sels, N)
    lo, hi = fp.box(N)
    rad = np.abs(R[-1].G).sum(axis=1)
    assert_close(lo, R[-1].c - rad, "box lo")
    M = np.zeros((2, n)); M[0, n - 1] = 1.0; M[1] = np.random.default_rng(2).standard_normal(n)
    Cp, Rp = fp.project(M)
    assert_close(Rp[:, N - 1], np.abs(M @ R[-1].G).sum(axis=1), "projected radius")
    fp.close()
    from reachability_jl_b200 import ReachError
    with pytest.raises(ReachError) as ei:               # the documented ceiling of this build
        ctx.glgm06(np.eye(4100), np.zeros(4100), np.eye(4100), 3, 2, np.zeros(4100), np.eye(4100), 0)
    assert ei.value.code == 3


def test_config_c2_full_size_properties(ctx):
    """BASELINE.json configs[1] at full size (n = 270, max_order 10, N = 2000, 2700 generators per reach set, 11.7 GB):
    (a) the first 120 reach sets -- growth phase, first reductions, steady state -- equal the oracle's entry by entry;
    (b) size-independent properties over all 2000 steps: generator counts, closed-form centres at sampled steps
        (c_k = Phi^{k-1} c0 + sum_{j<k-1} Phi^j cV, reduce_order never moves the centre), the box part of every
        sampled reach set is axis-aligned and non-negative, and a second run reproduces the first bit for bit."""
    from reachability_jl_b200 import workloads as wl
    w = wl.iss()
    Phi = orc.exp_Adelta(w.A, w.delta)
    P2 = orc.phi12_series(np.abs(w.A), w.delta)[2]
    D = orc.discretize_zonotope(w.A, orc.Box(w.c, w.r), w.delta, w.B, orc.Box(w.cU, w.rU), Phi=Phi, Phi2abs=P2)
    Om = orc.reduce_order(D.Omega0, w.max_order)
    N = orc.n_steps_glgm06(w.T, w.delta)
    n, r = w.n, w.max_order
    assert N == 2000 and Om.ngens == 1084 and D.V.ngens == 273
    fp = ctx.glgm06(D.Phi, Om.c, Om.G, N, r, D.V.c, D.V.G, 0).run()
    p = fp.ngens()
    Ro = orc.glgm06_reach_inhomog(Om, D.V, D.Phi, 120, r)
    for k in list(range(1, 16)) + [40, 77, 120]:
        c, G = fp.step(k, int(p[k - 1]))
        assert G.shape == Ro[k - 1].G.shape
        assert_close(c, Ro[k - 1].c, f"centre {k}")
        assert_close(G, Ro[k - 1].G, f"generators {k}")
    assert list(p[:7]) == [1084, 1357, 1630, 1903, 2176, 2449, 2700] and np.all(p[6:] == r * n)
    acc = np.zeros(n)           # sum_{j < k-1} Phi^j cV
    Pj = np.eye(n)
    samples = {2, 3, 500, 1000, 2000}
    keep = {}
    for k in range(1, N + 1):
        if k in samples:
            keep[k] = Pj @ Om.c + acc if k > 1 else Om.c.copy()      # Pj = Phi^{k-1}
        acc = acc + Pj @ D.V.c
        Pj = Pj @ D.Phi
    for k in sorted(samples):
        c, G = fp.step(k, int(p[k - 1]))
        assert_close(c, keep[k], f"closed-form centre {k}")
        if k >= 7:              # reduced sets end with the n box generators
            box = G[:, -n:]
            assert np.count_nonzero(box - np.diag(np.diag(box))) == 0 and np.all(np.diag(box) > 0)
        assert np.all(np.isfinite(G))
    # (c) device-side projection of all 2000 reach sets (project_reach.jl:18-99) against the fetched zonotopes and the box hull
    M = np.zeros((2, n)); M[0, 3] = 1.0; M[1] = np.random.default_rng(1).standard_normal(n)
    Cp, Rp = fp.project(M)
    for k in sorted(samples):
        c, G = fp.step(k, int(p[k - 1]))
        assert_close(Cp[:, k - 1], M @ c, f"projected centre {k}")
        assert_close(Rp[:, k - 1], np.abs(M @ G).sum(axis=1), f"projected radius {k}")
        lo, hi = fp.box(k)
        assert_close(Cp[0, k - 1] - Rp[0, k - 1], lo[3], f"coordinate projection = box hull {k}")
    assert np.all(np.isfinite(Cp)) and np.all(Rp > 0)
    c_a, G_a = fp.step(1999)
    fp.run()
    c_b, G_b = fp.step(1999)
    assert np.array_equal(c_a, c_b) and np.array_equal(G_a, G_b)      # deterministic: no float atomics, fixed orders
    fp.close()


# ---- projection of the zonotope flowpipe on the device (project_reach.jl:18-99) ------------------------------------
@pytest.mark.parametrize("n,m,N,r,q", [(9, 2, 14, 3, 2), (16, 1, 10, 2, 1), (31, 3, 12, 2, 4), (6, 0, 9, 5, 3)])
def test_projection_matches_host_linear_map(ctx, n, m, N, r, q):
    """`project(rs, M)` = linear_map(M, Z) per reach set, Hyperrectangle overapproximation of the result
    (set_type_proj default): device kernels vs numpy on the oracle's zonotopes."""
    A, X0, B, U, delta = _problem(300 + n, n, m)
    D, Om, R, _ = _oracle(A, X0, B, U, delta, N, r, True)
    fp = _device(ctx, D, Om, N, r, True, record=False)
    M = np.random.default_rng(n).standard_normal((q, n))
    if q >= 2:
        M[1] = 0.0
        M[1, n - 1] = 1.0           # a plain coordinate selection, as project_reach builds it
    C, Rad, Gp = fp.project(M, generators=True)
    p = fp.ngens()
    for k in range(1, N + 1):
        Z = R[k - 1]
        assert_close(C[:, k - 1], M @ Z.c, f"step {k} projected centre")
        MG = M @ Z.G
        assert_close(Gp[:, :p[k - 1], k - 1], MG, f"step {k} projected generators")
        assert np.all(Gp[:, p[k - 1]:, k - 1] == 0.0)
        assert_close(Rad[:, k - 1], np.abs(MG).sum(axis=1), f"step {k} projected box radius")
    C2, Rad2 = fp.project(M)                   # boxes only: identical bits (fixed summation order)
    assert np.array_equal(C2, C) and np.array_equal(Rad2, Rad)
    fp.close()


def test_projection_rejects_bad_arguments(ctx):
    A, X0, B, U, delta = _problem(7, 5, 1)
    D, Om, R, _ = _oracle(A, X0, B, U, delta, 4, 3, True)
    fp = _device(ctx, D, Om, 4, 3, True, record=False)
    from reachability_jl_b200 import ReachError
    with pytest.raises(ReachError):
        fp.project(np.ones((5, 5)))            # q > 4
    with pytest.raises(ReachError):
        fp.project(np.ones((2, 5)), Gp=np.zeros((2, 1, 4), order="F"))     # pmax too small
    fp.close()


# ---- structurally sparse Phi: the generator map as an ELL sparse product ----------------------------------------------
@pytest.mark.parametrize("n,m,N,r", [(40, 2, 45, 3), (33, 1, 70, 2)])
def test_sparse_phi_map_equals_dense_gemm(ctx, n, m, N, r, monkeypatch):
    """Decoupled 2 x 2 modes: Phi and its powers have two non-zeros per row, so linear_map(Phi^w, .) runs as a sparse
    product.  Same results as the DMMA GEMM (forced with REACH_GLGM06_DENSE) -- the skipped terms are exact zeros -- and both
    agree with the oracle, order-reduction selections included."""
    g = np.random.default_rng(n)
    A, X0, B, U, delta = _problem(500 + n, n, m, dense=False)
    Phi = np.zeros((n, n))                       # exact block-diagonal exponential (scipy's Pade would fill in 1e-20 noise)
    for i in range(n // 2):
        blk = A[2 * i:2 * i + 2, 2 * i:2 * i + 2]
        Phi[2 * i:2 * i + 2, 2 * i:2 * i + 2] = orc.exp_Adelta(blk, delta)
    if n % 2:
        Phi[-1, -1] = np.exp(A[-1, -1] * delta)
    D = orc.discretize_zonotope(A, X0, delta, B, U, rzg=True, Phi=Phi)
    Om = orc.reduce_order(D.Omega0, r, True)
    R, sels = orc.glgm06_reach_inhomog(Om, D.V, D.Phi, N, r, True, return_selections=True)
    fp = _device(ctx, D, Om, N, r, True)
    info = fp.map_info()
    assert info["sparse_levels"] >= 1 and info["nnz_per_row"] == 2
    _compare(fp, R, sels, N)
    monkeypatch.setenv("REACH_GLGM06_DENSE", "1")
    fd = _device(ctx, D, Om, N, r, True)
    assert fd.map_info()["sparse_levels"] == 0
    p = fp.ngens()
    assert np.array_equal(p, fd.ngens())
    for k in (1, 2, N // 2, N):
        cs, Gs = fp.step(k, int(p[k - 1]))
        cd, Gd = fd.step(k, int(p[k - 1]))
        assert_close(cs, cd, f"step {k} centre: sparse map vs dense GEMM")
        assert_close(Gs, Gd, f"step {k} generators: sparse map vs dense GEMM")
    fp.close(); fd.close()


def test_dense_phi_keeps_the_gemm(ctx):
    A, X0, B, U, delta = _problem(3, 24, 2)
    D, Om, R, sels = _oracle(A, X0, B, U, delta, 6, 3, True)
    fp = _device(ctx, D, Om, 6, 3, True)
    assert fp.map_info() == dict(sparse_levels=0, nnz_per_row=0)
    fp.close()


def test_config_c5_ensemble_full_length_properties(ctx):
    """BASELINE.json configs[4] at one GPU's share of the 4096 splits (256 of the 512 members here, to bound the test's 10 GB):
    heli28-shaped n = 28, N = 2000 reach sets per member.  (a) sampled members equal their own oracle flowpipe at sampled
    steps up to the last one; (b) closed form R_k = Phi^{k-1} (c0, G0) for every member at the last step through the box
    hull; (c) the boxes of the sub-boxes stay inside the box of the parent set's flowpipe (monotonicity of the split)."""
    from reachability_jl_b200 import workloads as wl
    w = wl.heli28()
    n, N = w.n, orc.n_steps_glgm06(w.T, w.delta)
    assert N == 2000
    Phi = orc.exp_Adelta(w.A, w.delta)
    P2 = orc.Phi2(np.abs(w.A), w.delta)
    cs, rs = wl.split_box(w.c, w.r, [4, 4, 4, 4, 2, 2, 2, 2])
    assert cs.shape[1] == 4096
    batch = 256
    members = np.linspace(0, 4095, batch).astype(int)
    zs = [orc.discretize_zonotope(w.A, orc.Box(cs[:, b], rs[:, b]), w.delta, rzg=False, Phi=Phi, Phi2abs=P2).Omega0 for b in members]
    p = zs[0].ngens
    assert p == 85 and all(z.ngens == p for z in zs)
    c0s = np.stack([z.c for z in zs], axis=1)
    G0s = np.stack([z.G for z in zs], axis=2)
    ens = ctx.ensemble(Phi, c0s, G0s, N).run()
    for b in (0, 97, batch - 1):
        R = orc.glgm06_reach_homog(zs[b], Phi, N, rzg=False)
        for k in (1, 2, 500, N):
            c, G = ens.step(b, k)
            assert_close(c, R[k - 1].c, f"member {b} step {k} centre")
            assert_close(G, R[k - 1].G, f"member {b} step {k} generators")
    PN = np.linalg.matrix_power(Phi, N - 1)
    lo, hi = ens.boxes(N)
    cN = PN @ c0s
    radN = np.stack([np.abs(PN @ G0s[:, :, b]).sum(axis=1) for b in range(batch)], axis=1)
    scale = np.abs(cN) + radN
    assert np.max(np.abs(lo - (cN - radN)) / (1e-9 * scale + 1e-12)) <= 1.0       # matrix_power takes another product order
    assert np.max(np.abs(hi - (cN + radN)) / (1e-9 * scale + 1e-12)) <= 1.0
    parent = orc.discretize_zonotope(w.A, orc.Box(w.c, w.r), w.delta, rzg=False, Phi=Phi, Phi2abs=P2).Omega0
    Rp = orc.glgm06_reach_homog(parent, Phi, 40, rzg=False)
    lo40, hi40 = ens.boxes(40)
    rad = np.abs(Rp[39].G).sum(axis=1)
    tol = 1e-9 * (np.abs(Rp[39].c) + rad) + 1e-12
    assert np.all(lo40 >= (Rp[39].c - rad - tol)[:, None]) and np.all(hi40 <= (Rp[39].c + rad + tol)[:, None])
    ens.close()
