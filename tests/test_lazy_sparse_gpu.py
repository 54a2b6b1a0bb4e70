"""GPU parity: the sparse-A / lazy-expm variants (`reach_sparse_*`, `reach_expmv`, `reach_discretize_box_sparse`,
`reach_bffpsv18_boxes_lazy`) against the oracle.

Reference behaviour: BFFPSV18 with `:lazy_expm` (BFFPSV18/reach_blocks.jl:227-397, rows of exp(kδA) through
`get_rows`) and `discretize` with a lazy exponential (discretize.jl:153-154, 302-306).  The oracle evaluates the same
quantities with dense matrix exponentials (`bffpsv18_reach`, `bffpsv18_reach_lazy` literal restatement, `discretize_box`);
tolerance 1e-10 relative + 1e-12 absolute as everywhere (tests/util.py)."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import reach_oracle as orc
from tests.util import assert_close

pytestmark = pytest.mark.gpu


def _sparse_system(seed, n, density=0.05, stiff=3.0):
    g = np.random.default_rng(seed)
    M = sp.random(n, n, density=density, random_state=np.random.RandomState(seed), format="csc")
    M.data = g.standard_normal(len(M.data))
    A = M.toarray()
    np.fill_diagonal(A, 0.0)
    A -= np.diag(g.uniform(0.5, stiff, n) + np.abs(A).sum(axis=1))      # diagonally dominant, stable
    return A


@pytest.mark.parametrize("n,R,delta,transpose", [(7, 1, 0.1, False), (40, 5, 0.05, True), (123, 33, 0.02, False),
                                                 (64, 64, 0.7, True)])
def test_expmv_matches_dense_expm(ctx, n, R, delta, transpose):
    A = _sparse_system(n, n)
    X = np.random.default_rng(n + 1).standard_normal((n, R))
    S = ctx.sparse(sp.csc_matrix(A))
    Y = S.expmv(delta, X, transpose=transpose)
    Phi = orc.exp_Adelta(A, delta)
    assert_close(Y, (Phi.T if transpose else Phi) @ X, "expmv")
    S.close()


def test_sparse_create_accepts_dense_and_rejects_bad_csc(ctx):
    from reachability_jl_b200 import ReachError
    import ctypes as C
    A = _sparse_system(3, 12)
    S1, S2 = ctx.sparse(A), ctx.sparse(sp.csc_matrix(A))
    x = np.arange(12.0)
    assert np.array_equal(S1.expmv(0.1, x), S2.expmv(0.1, x))
    S1.close(); S2.close()
    colptr = np.array([1, 2, 4], dtype=np.int64)         # colptr[n] != nnz + 1
    rowval = np.array([1, 2], dtype=np.int64)
    val = np.ones(2)
    h = C.c_void_p()
    i64p = C.POINTER(C.c_int64)
    rc = ctx.lib.reach_sparse_create(ctx._h, 2, 2, colptr.ctypes.data_as(i64p), rowval.ctypes.data_as(i64p),
                                     val.ctypes.data_as(C.POINTER(C.c_double)), C.byref(h))
    assert rc == 1 and not h.value


@pytest.mark.parametrize("n,m,algorithm,sparse_r", [(30, 0, "forward", False), (45, 2, "forward", True),
                                                    (45, 2, "backward", False), (18, 18, "forward", False)])
def test_sparse_discretize_matches_dense_oracle(ctx, n, m, algorithm, sparse_r):
    g = np.random.default_rng(100 + n + m)
    A = _sparse_system(n, n)
    c = g.standard_normal(n)
    r = np.abs(g.standard_normal(n)) * 0.1
    if sparse_r:
        r[n // 3:] = 0.0                                  # only some columns of Φ matter for |Φ| r
    B = U = None
    if m:
        B = None if m == n else g.standard_normal((n, m))
        U = orc.Box(g.standard_normal(m), np.abs(g.standard_normal(m)) * 0.2)
    D = orc.discretize_box(A, orc.Box(c, r), 0.02, B, U, algorithm=algorithm)
    S = ctx.sparse(sp.csc_matrix(A))
    d = S.discretize_box(0.02, c, r, B, None if U is None else U.c, None if U is None else U.r, algorithm=algorithm)
    assert_close(d["c0"], D.c0, "c0")
    assert_close(d["r0"], D.r0, "r0")
    assert_close(d["rE"], D.rE, "rE")
    if m:
        assert_close(d["rpsi"], D.rpsi, "rpsi")
        assert_close(d["MU"], D.MU, "MU")
    S.close()


def test_sparse_discretize_rejects_other_models(ctx):
    from reachability_jl_b200 import ReachError
    S = ctx.sparse(_sparse_system(1, 6))
    with pytest.raises(ReachError) as ei:
        S.discretize_box(0.01, np.zeros(6), np.ones(6), algorithm="firstorder")
    assert ei.value.code == 3
    with pytest.raises(ValueError):
        S.discretize_box(0.01, np.zeros(6), np.ones(6), algorithm="sideways")
    S.close()


@pytest.mark.parametrize("n,m,N,rows", [(24, 0, 12, None), (37, 2, 25, [0, 5, 36]), (80, 3, 40, list(range(10, 50))),
                                        (16, 1, 1, [3])])
def test_lazy_boxes_match_dense_recurrence_and_literal_rows(ctx, n, m, N, rows):
    g = np.random.default_rng(7 * n + m)
    A = _sparse_system(2 * n + 1, n)
    X0 = orc.Box(g.standard_normal(n), np.abs(g.standard_normal(n)) * 0.1)
    B = g.standard_normal((n, m)) if m else None
    U = orc.Box(g.standard_normal(m), np.abs(g.standard_normal(m)) * 0.2) if m else None
    delta = 0.02
    D = orc.discretize_box(A, X0, delta, B, U)
    Co, Ro = orc.bffpsv18_reach(D, N, rows)
    S = ctx.sparse(sp.csc_matrix(A))
    Cd, Rd = S.bffpsv18_boxes(delta, D.c0, D.r0, N, D.MU, D.cU, D.rU, D.rpsi,
                              None if rows is None else np.asarray(rows, dtype=np.int32))
    assert_close(Cd.T, Co, "lazy centres vs dense recurrence")
    assert_close(Rd.T, Ro, "lazy radii vs dense recurrence")
    if N <= 25:       # the reference's own formulation: rows of exp(kδA) from scratch at every step
        Cl, Rl = orc.bffpsv18_reach_lazy(A, D, delta, N, rows)
        assert_close(Cd.T, Cl, "lazy centres vs from-scratch rows")
        assert_close(Rd.T, Rl, "lazy radii vs from-scratch rows")
    S.close()


def test_lazy_path_end_to_end_matches_dense_device_path(ctx):
    """Sparse discretisation + lazy rows against the dense device path (expm / Φ₂ GEMM chain + stacked DMMA power
    GEMM) on a C4-shaped system at reduced size: two independent device implementations of the same flowpipe."""
    from reachability_jl_b200 import workloads as wl
    w = wl.mna5(n=600, N=60, seed=5)
    rows = np.array([0, 1, 2, 9, 77, 333, 599], dtype=np.int32)
    N = 60
    Dd = ctx.discretize_box(w.A, w.delta, w.c, w.r, w.B, w.cU, w.rU)
    C1, R1 = ctx.bffpsv18_boxes(Dd["Phi"], Dd["c0"], Dd["r0"], N, Dd["MU"], Dd["cU"], Dd["rU"], Dd["rpsi"], rows)
    S = ctx.sparse(sp.csc_matrix(w.A))
    Ds = S.discretize_box(w.delta, w.c, w.r, w.B, w.cU, w.rU)
    for key in ("c0", "r0", "rE", "rpsi", "MU"):
        assert_close(Ds[key], Dd[key], key)
    C2, R2 = S.bffpsv18_boxes(w.delta, Ds["c0"], Ds["r0"], N, Ds["MU"], Ds["cU"], Ds["rU"], Ds["rpsi"], rows)
    assert_close(C2, C1, "centres")
    assert_close(R2, R1, "radii")
    t = S.timing()
    assert t["products"] > 0 and t["run_ms"] > 0
    S.close()


def test_graph_replay_and_eager_steps_produce_identical_bits(ctx, monkeypatch):
    """Two steps of the lazy recurrence are captured into a CUDA graph and replayed (output column from a device counter);
    REACH_LAZY_NO_GRAPH issues the same kernels eagerly (what the ncu launch lists profile)."""
    g = np.random.default_rng(77)
    n, m, N = 150, 3, 34
    A = _sparse_system(9, n, density=0.03, stiff=8.0)         # ‖δA‖ > 2: more than one scaling repetition
    X0 = orc.Box(g.standard_normal(n), np.abs(g.standard_normal(n)) * 0.1)
    B = g.standard_normal((n, m))
    U = orc.Box(g.standard_normal(m), np.abs(g.standard_normal(m)) * 0.2)
    delta = 0.2
    D = orc.discretize_box(A, X0, delta, B, U)
    rows = np.arange(0, n, 3, dtype=np.int32)
    S = ctx.sparse(sp.csc_matrix(A))
    C1, R1 = S.bffpsv18_boxes(delta, D.c0, D.r0, N, D.MU, D.cU, D.rU, D.rpsi, rows)
    monkeypatch.setenv("REACH_LAZY_NO_GRAPH", "1")
    C2, R2 = S.bffpsv18_boxes(delta, D.c0, D.r0, N, D.MU, D.cU, D.rU, D.rpsi, rows)
    assert np.array_equal(C1, C2) and np.array_equal(R1, R2)
    # rows of interest are processed in L2-sized panels; every row is independent, so the panel width cannot change a bit
    monkeypatch.setenv("REACH_LAZY_PANEL", "12")
    C3, R3 = S.bffpsv18_boxes(delta, D.c0, D.r0, N, D.MU, D.cU, D.rU, D.rpsi, rows)
    assert np.array_equal(C1, C3) and np.array_equal(R1, R3)
    Co, Ro = orc.bffpsv18_reach(D, N, rows)
    assert_close(C1.T, Co, "centres")
    assert_close(R1.T, Ro, "radii")
    S.close()


@pytest.mark.parametrize("n,m,N,q", [(30, 0, 15, 1), (42, 2, 22, 2), (25, 3, 18, 3)])
def test_lazy_output_map_and_check_match_the_dense_oracle(ctx, n, m, N, q):
    """output_function branch and check mode with a lazy exponential: blocks stay lazy, so the absolute value is taken per
    block of the product M[:, b_i] P[b_i, :] (oracle `bffpsv18_output_map` / `bffpsv18_check`)."""
    g = np.random.default_rng(31 * n + m)
    A = _sparse_system(3 * n + 2, n)
    X0 = orc.Box(g.standard_normal(n), np.abs(g.standard_normal(n)) * 0.1)
    B = g.standard_normal((n, m)) if m else None
    U = orc.Box(g.standard_normal(m), np.abs(g.standard_normal(m)) * 0.2) if m else None
    delta = 0.02
    D = orc.discretize_box(A, X0, delta, B, U)
    rows = np.array([0, 1, 4, 5, 6, 11, 12, n - 1], dtype=np.int32)
    sizes = [2, 3, 2, 1]                                  # blocks of the partition the rows belong to
    M = g.standard_normal((q, len(rows)))
    Co, Ro = orc.bffpsv18_output_map(D, N, M, rows, sizes)
    S = ctx.sparse(sp.csc_matrix(A))
    Cd, Rd = S.bffpsv18_output_map(delta, D.c0, D.r0, N, M, D.MU, D.cU, D.rU, D.rpsi, rows, block_sizes=sizes)
    assert_close(Cd.T, Co, "output map centres")
    assert_close(Rd.T, Ro, "output map radii")
    # check mode on the first output row: thresholds around the largest support value
    ell = M[0]
    rho = Co[:, 0] + Ro[:, 0]
    for bound in (rho.max() + 1.0, np.sort(rho)[len(rho) // 2] * (1 - 1e-6) - 1e-9, rho.min() - 1.0):
        want = orc.bffpsv18_check(D, N, ell, bound, rows, eager=True, block_sizes=sizes)
        got = S.bffpsv18_check(delta, D.c0, D.r0, N, ell, bound, D.MU, D.cU, D.rU, D.rpsi, rows, block_sizes=sizes)
        assert got == want, (bound, got, want)
    S.close()


@pytest.mark.parametrize("name", ["lcs4", "clccs4", "doc5"])
def test_lazy_path_reproduces_the_golden_fixtures(ctx, name):
    """The committed golden vectors (tests/golden/*.npz: the reference's own test systems, frozen oracle outputs) through the
    sparse discretisation and the lazy recurrence -- no oracle call at run time."""
    from pathlib import Path
    Gd = np.load(Path(__file__).resolve().parent / "golden" / f"{name}.npz")
    A, c, r, delta, T = Gd["A"], Gd["c"], Gd["r"], float(Gd["delta"]), float(Gd["T"])
    B = Gd["B"] if "B" in Gd else None
    cU = Gd["cU"] if "cU" in Gd else None
    rU = Gd["rU"] if "rU" in Gd else None
    S = ctx.sparse(sp.csc_matrix(A))
    for alg in ("forward", "backward"):
        D = S.discretize_box(delta, c, r, B, cU, rU, algorithm=alg)
        assert_close(D["c0"], Gd[f"{alg}_c0"], f"{alg} c0")
        assert_close(D["r0"], Gd[f"{alg}_r0"], f"{alg} r0")
        assert_close(D["rE"], Gd[f"{alg}_rE"], f"{alg} rE")
        if B is not None:
            assert_close(D["rpsi"], Gd[f"{alg}_rpsi"], f"{alg} rpsi")
        N = Gd[f"{alg}_centers"].shape[0]
        C, R = S.bffpsv18_boxes(delta, D["c0"], D["r0"], N, D["MU"], D["cU"], D["rU"], D["rpsi"])
        assert_close(C.T, Gd[f"{alg}_centers"], f"{alg} centres")
        assert_close(R.T, Gd[f"{alg}_radii"], f"{alg} radii")
    S.close()
