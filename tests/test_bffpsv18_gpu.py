"""GPU parity: BFFPSV18 device path (through the C ABI) vs the CPU oracle.

Tolerance (north_star): 1e-10 relative + 1e-12 absolute on every box centre and radius.
Mirrors the shapes of /root/reference/test/Reachability/solve_continuous.jl (4x4 LCS, affine CLCCS, `:vars` subsets,
1-D Interval and 2-D Hyperrectangle partitions) plus BASELINE.json's configs C1 and C3.
"""
from pathlib import Path

import numpy as np
import pytest

from oracle import reach_oracle as orc
from tests.util import assert_close

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"


def _run(ctx, D, N, rows=None):
    C, R = ctx.bffpsv18_boxes(D.Phi, D.c0, D.r0, N, D.MU, D.cU, D.rU, D.rpsi, rows)
    return C.T, R.T       # (N, nrows) like the oracle


def _check(ctx, D, N, rows=None):
    Co, Ro = orc.bffpsv18_reach(D, N, rows)
    Cg, Rg = _run(ctx, D, N, rows)
    assert_close(Cg, Co, "centres")
    assert_close(Rg, Ro, "radii")
    return Cg, Rg


@pytest.mark.parametrize("name", ["lcs4", "clccs4", "doc5"])
@pytest.mark.parametrize("alg", ["forward", "backward"])
def test_golden_fixtures(ctx, name, alg):
    z = np.load(GOLDEN / f"{name}.npz")
    N = z[f"{alg}_centers"].shape[0]
    inh = "B" in z
    C, R = ctx.bffpsv18_boxes(z[f"{alg}_Phi"], z[f"{alg}_c0"], z[f"{alg}_r0"], N,
                              z[f"{alg}_MU"] if inh else None, z["cU"] if inh else None, z["rU"] if inh else None,
                              z[f"{alg}_rpsi"] if inh else None)
    assert_close(C.T, z[f"{alg}_centers"], "centres")
    assert_close(R.T, z[f"{alg}_radii"], "radii")


def _random_problem(seed, n, m):
    g = np.random.default_rng(seed)
    A = g.standard_normal((n, n)) / np.sqrt(n) - 1.2 * np.eye(n)
    X0 = orc.Box(g.standard_normal(n), np.abs(g.standard_normal(n)) * 0.1)
    B = g.standard_normal((n, m)) if m else None
    U = orc.Box(g.standard_normal(m), np.abs(g.standard_normal(m)) * 0.3) if m else None
    return orc.discretize_box(A, X0, 0.01, B, U)


@pytest.mark.parametrize("n,m,N", [(1, 0, 5), (1, 1, 40), (2, 1, 3), (7, 3, 33), (8, 0, 64), (9, 2, 65), (33, 5, 129),
                                   (130, 4, 257), (257, 1, 100), (300, 17, 40)])
def test_random_systems_ragged_sizes(ctx, n, m, N):
    _check(ctx, _random_problem(100 + n, n, m), N)


@pytest.mark.parametrize("depth", [1, 2, 3, 5, 7, 12, 49, 64])
@pytest.mark.parametrize("n,m,N", [(40, 2, 131), (130, 0, 101)])
def test_any_stack_depth(ctx, n, m, N, depth, monkeypatch):
    """The stacked power GEMM advances `depth` steps per launch; any depth works, not only powers of two (the first stack is
    built by doubling with a partial last round, Phi^depth is the product of the squares at the set bits of depth), and with
    more than one block the radii come out of the GEMM's fused abs-sum epilogue.  Same sets for every depth."""
    monkeypatch.setenv("REACH_BFFPSV18_BLOCK", str(depth))
    D = _random_problem(600 + n, n, m)
    _check(ctx, D, N)
    _check(ctx, D, N, rows=np.array([n - 1, 0, 3], dtype=np.int32))


def test_fused_abs_sum_epilogue_equals_streaming_pass(ctx, monkeypatch):
    """|P_k| r0 and |P_k| rpsi from the stacked GEMM's accumulators (fused epilogue, per-column-group partials) against the
    separate streaming pass over the stack (REACH_BFFPSV18_NO_FUSED_ABS=1): same sets to rounding, each bit-reproducible."""
    D = _random_problem(77, 200, 3)
    N = 150
    monkeypatch.setenv("REACH_BFFPSV18_BLOCK", "6")
    Cf, Rf = _run(ctx, D, N)
    Cf2, Rf2 = _run(ctx, D, N)
    assert np.array_equal(Cf, Cf2) and np.array_equal(Rf, Rf2)
    monkeypatch.setenv("REACH_BFFPSV18_NO_FUSED_ABS", "1")
    Cu, Ru = _run(ctx, D, N)
    assert np.array_equal(Cf, Cu)                      # centres never involve the abs-sums
    assert_close(Rf, Ru, "radii: fused epilogue vs streaming pass")
    Co, Ro = orc.bffpsv18_reach(D, N)
    assert_close(Rf, Ro, "radii vs oracle")


@pytest.mark.parametrize("N", [1, 2, 3])
def test_tiny_step_counts(ctx, N):
    """N == 1 returns only X_1 = Xhat0[blocks] (reach.jl:87-90, reach_blocks.jl:152-158)."""
    D = _random_problem(5, 12, 2)
    C, R = _check(ctx, D, N)
    assert np.array_equal(C[0], D.c0) and np.array_equal(R[0], D.r0)


def test_rows_subset_and_order(ctx):
    """`:vars` -> blocks -> dimensions (compute_dimensions.jl:1-15): any subset, any order, no duplicates needed."""
    D = _random_problem(6, 40, 3)
    _check(ctx, D, 50, rows=[3])
    _check(ctx, D, 50, rows=[0, 1, 38, 39])
    _check(ctx, D, 50, rows=[39, 5, 17, 4, 20, 21, 22, 23, 24])
    part = [[2 * i, 2 * i + 1] for i in range(20)]
    blocks = orc.compute_blocks([1, 3, 40], [[a + 1 for a in b] for b in part])
    rows = [d - 1 for d in orc.compute_dimensions([[a + 1 for a in b] for b in part], blocks)]
    assert rows == [0, 1, 2, 3, 38, 39]
    _check(ctx, D, 50, rows=rows)


def test_zero_radius_and_singleton_inputs(ctx):
    """Flat initial sets (r = 0) and a Singleton input (rU = 0): affine term x' = Ax + b (solve_continuous.jl:216-224)."""
    g = np.random.default_rng(3)
    n = 10
    A = g.standard_normal((n, n)) - 2 * np.eye(n)
    D = orc.discretize_box(A, orc.Box(g.standard_normal(n), np.zeros(n)), 0.02, np.eye(n),
                           orc.Box(g.standard_normal(n), np.zeros(n)))
    _check(ctx, D, 30)


def test_invalid_arguments_raise(ctx):
    import reachability_jl_b200 as rb
    D = _random_problem(5, 6, 0)
    with pytest.raises(rb.ReachError) as ei:
        ctx.bffpsv18_boxes(D.Phi, D.c0, D.r0, 5, rows=[0, 6])
    assert ei.value.code == 1 and "row index" in str(ei.value)
    with pytest.raises(rb.ReachError):
        ctx.bffpsv18_boxes(D.Phi, D.c0, D.r0, 0)


def _workload_problem(w):
    X0 = orc.Box(w.c, w.r)
    U = orc.Box(w.cU, w.rU) if w.B is not None else None
    Phi = orc.exp_Adelta(w.A, w.delta)
    P2 = orc.phi12_series(np.abs(w.A), w.delta)[2]      # = Φ₂(|A|, δ) (tests/test_oracle_pins.py); avoids the 3n x 3n exp
    return orc.discretize_box(w.A, X0, w.delta, w.B, U, Phi=Phi, Phi2abs=P2)


def test_config_c1_building_full(ctx):
    """BASELINE.json configs[0]: n = 48, 1-D Interval blocks, δ = 0.003, T = 20 -> N = 6667, all variables."""
    from reachability_jl_b200 import workloads as wl
    w = wl.building()
    D = _workload_problem(w)
    N = orc.n_steps_bffpsv18(w.T, w.delta)
    assert N == 6667
    _check(ctx, D, N)


def test_config_c3_fom_prefix_and_rows(ctx):
    """BASELINE.json configs[2]: n = 1006, 503 2-D Hyperrectangle blocks; oracle-sized prefix N = 150."""
    from reachability_jl_b200 import workloads as wl
    w = wl.fom()
    D = _workload_problem(w)
    _check(ctx, D, 150)
    _check(ctx, D, 40, rows=list(range(100, 226)))


def test_config_c3_full_size_properties(ctx):
    """Full N = 6667 at n = 1006: size-independent properties instead of the (hours-long) CPU loop.
    (a) prefix: the first 150 steps equal the oracle's; (b) closed form at sampled steps k:
        homogeneous run  c_k = Φ^{k-1} c0,  r_k = |Φ^{k-1}| r0  with Φ^{k-1} by repeated squaring on the CPU;
    (c) linearity: doubling (c0, r0, cU, rU, rψ) doubles every bound; (d) radii are non-negative, and the
        inhomogeneous radii dominate the homogeneous ones."""
    from reachability_jl_b200 import workloads as wl
    w = wl.fom()
    D = _workload_problem(w)
    N = orc.n_steps_bffpsv18(w.T, w.delta)
    Cg, Rg = _run(ctx, D, N)
    Co, Ro = orc.bffpsv18_reach(D, 150)
    assert_close(Cg[:150], Co, "prefix centres")
    assert_close(Rg[:150], Ro, "prefix radii")
    Dh = orc.DiscreteBoxProblem(D.Phi, D.c0, D.r0, None, None, None, None, D.rE, D.Phi2abs)
    Ch, Rh = _run(ctx, Dh, N)
    for k in (2, 500, 3333, 6667):
        Pk = np.linalg.matrix_power(D.Phi, k - 1)
        assert_close(Ch[k - 1], Pk @ D.c0, f"closed-form centre {k}")
        assert_close(Rh[k - 1], np.abs(Pk) @ D.r0, f"closed-form radius {k}")
    assert np.all(Rg >= 0) and np.all(Rg >= Rh * (1 - 1e-12))
    D2 = orc.DiscreteBoxProblem(D.Phi, 2 * D.c0, 2 * D.r0, D.MU, 2 * D.cU, 2 * D.rU, 2 * D.rpsi, D.rE, D.Phi2abs)
    C2, R2 = _run(ctx, D2, N)
    assert_close(C2, 2 * Cg, "linearity centres")
    assert_close(R2, 2 * Rg, "linearity radii")


def test_boxplan_is_repeatable_and_counts_launches(ctx):
    D = _random_problem(9, 64, 2)
    before = ctx.launch_count()
    plan = ctx.boxplan(D.Phi, D.c0, D.r0, 200, D.MU, D.cU, D.rU, D.rpsi)
    plan.run()
    a = plan.fetch()
    plan.run()
    b = plan.fetch()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])      # deterministic summation order
    assert ctx.launch_count() > before
    t = plan.timing()
    assert t["run_ms"] > 0 and t["gemm_launches"] > 0
    plan.close()


def test_config_c4_mna5_full_dimension(ctx):
    """BASELINE.json configs[3]: n = 10913 (MNA5-shaped), 1-D blocks, delta = 1e-3.  Full state dimension on the GPU
    (Phi = 953 MB, 2.6 TFLOP per step); the oracle follows a handful of rows (its cost is O(rows n^2) per step) with
    the same Phi, which the device discretisation produces (device expm is pinned against the oracle at smaller n in
    tests/test_discretize_gpu.py).  Also checks the row sharding of SURVEY.md section 8e at this size: two disjoint
    row shards reproduce the rows of the unsharded run bit for bit."""
    from reachability_jl_b200 import workloads as wl
    from reachability_jl_b200 import sharding
    w = wl.mna5()
    n = w.n
    out = ctx.discretize_box(w.A, w.delta, w.c, w.r, w.B, w.cU, w.rU)
    D = orc.DiscreteBoxProblem(out["Phi"], out["c0"], out["r0"], out["MU"], out["cU"], out["rU"], out["rpsi"], out["rE"], None)
    # the device discretisation at this size: a few columns of Phi against the action of the sparse exponential
    # (scipy expm_multiply, Al-Mohy & Higham), and Omega0's box contains X0
    import scipy.sparse as sp
    from scipy.sparse.linalg import expm_multiply
    cols = [0, 5, n // 2, n - 1]
    I4 = np.zeros((n, len(cols)))
    I4[cols, range(len(cols))] = 1.0
    E = expm_multiply(sp.csc_matrix(w.A) * w.delta, I4)
    assert np.max(np.abs(out["Phi"][:, cols] - E)) < 1e-11
    assert np.all(out["c0"] - out["r0"] <= w.c - w.r + 1e-15) and np.all(out["c0"] + out["r0"] >= w.c + w.r - 1e-15)
    N = 12
    rows = np.array([0, 1, 9, 10, 1212, 5000, 5001, 10912], dtype=np.int32)
    Co, Ro = orc.bffpsv18_reach(D, N, rows)
    C, R = ctx.bffpsv18_boxes(D.Phi, D.c0, D.r0, N, D.MU, D.cU, D.rU, D.rpsi, rows)
    assert_close(C.T, Co, "C4 centres (row subset)")
    assert_close(R.T, Ro, "C4 radii (row subset)")
    # all rows, sharded over 2 "ranks" (contiguous whole blocks), vs the same rows taken from the subset run
    part = [[i] for i in range(1, n + 1)]
    blocks = list(range(1, n + 1))
    r0_ = sharding.shard_rows(part, blocks, 0, 2)
    r1_ = sharding.shard_rows(part, blocks, 1, 2)
    assert len(r0_) + len(r1_) == n and r0_[-1] + 1 == r1_[0]
    Cs0, Rs0 = ctx.bffpsv18_boxes(D.Phi, D.c0, D.r0, 4, D.MU, D.cU, D.rU, D.rpsi, r0_)
    Cs1, Rs1 = ctx.bffpsv18_boxes(D.Phi, D.c0, D.r0, 4, D.MU, D.cU, D.rU, D.rpsi, r1_)
    full_c = np.vstack([Cs0, Cs1])
    full_r = np.vstack([Rs0, Rs1])
    assert_close(full_c[rows, :], Co[:4].T, "sharded centres")
    assert_close(full_r[rows, :], Ro[:4].T, "sharded radii")
    assert np.all(full_r >= 0)


def _block_diagonal_phi(g, n, max_block, permute=True):
    """Φ that is block-diagonal up to a random permutation of the states: blocks of 1 .. max_block states."""
    Phi = np.zeros((n, n))
    i = 0
    sizes = []
    while i < n:
        w = int(min(n - i, g.integers(1, max_block + 1)))
        Phi[i:i + w, i:i + w] = g.standard_normal((w, w)) * (0.9 / np.sqrt(w))
        sizes.append(w)
        i += w
    if permute:
        perm = g.permutation(n)
        Phi = Phi[np.ix_(perm, perm)]
    return Phi, max(sizes)


@pytest.mark.parametrize("dense_csc", [True, False])
def test_sparse_csc_phi_is_bit_identical_to_dense(ctx, dense_csc, monkeypatch):
    """reach_blocks!(ϕ::SparseMatrixCSC, ...) (reach_blocks.jl:47-132): same sets as the dense method.  Path (b) of the sparse
    twin scatters the CSC fields into the dense layout, so the results equal those of Matrix(Φ) bit for bit: taken by a Φ whose
    sparsity graph is connected, and forced for a block-diagonal one with REACH_BFFPSV18_CSC_DENSE=1."""
    import scipy.sparse as sp
    g = np.random.default_rng(5)
    n, N = 41, 30
    Phi = np.zeros((n, n))
    for i in range(0, n - 1, 2):                      # block-diagonal Φ stays sparse under powers
        Phi[i:i + 2, i:i + 2] = g.standard_normal((2, 2)) * 0.6
    Phi[-1, -1] = 0.5
    if dense_csc:
        monkeypatch.setenv("REACH_BFFPSV18_CSC_DENSE", "1")
    else:
        Phi[np.arange(n - 1), np.arange(1, n)] += 0.01           # a chain couples everything: one component of 41 states
    c0, r0 = g.standard_normal(n), np.abs(g.standard_normal(n))
    MU = g.standard_normal((n, 2)) * 0.1
    cU, rU, rpsi = g.standard_normal(2), np.abs(g.standard_normal(2)), np.abs(g.standard_normal(n)) * 0.01
    rows = np.array([0, 1, 7, 40], dtype=np.int32)
    Cd, Rd = ctx.bffpsv18_boxes(Phi, c0, r0, N, MU, cU, rU, rpsi, rows)
    Cs, Rs = ctx.bffpsv18_boxes_csc(sp.csc_matrix(Phi), c0, r0, N, MU, cU, rU, rpsi, rows)
    assert not ctx.csc_info()["sparse_path"]
    assert np.array_equal(Cs, Cd) and np.array_equal(Rs, Rd)
    P = orc.DiscreteBoxProblem(Phi, c0, r0, MU, cU, rU, rpsi, None, None)
    Co, Ro = orc.bffpsv18_reach(P, N, rows)
    assert_close(Cs.T, Co, "centres")
    assert_close(Rs.T, Ro, "radii")


@pytest.mark.parametrize("n,max_block,m,N,subset", [(41, 2, 2, 60, True), (97, 4, 0, 130, False), (150, 8, 3, 90, True),
                                                    (203, 16, 1, 75, False), (64, 1, 2, 40, False)])
def test_block_diagonal_csc_phi_skips_the_zero_blocks(ctx, n, max_block, m, N, subset):
    """Path (a) of the sparse twin (reach_blocks.jl:47-132 skips zero blocks): decoupled subsystems scattered by a random
    permutation of the states.  Every row of Φ^k lives on its component, the recurrence costs w^2 fma per row and step; the sets
    equal the oracle's and the dense DMMA path's."""
    import scipy.sparse as sp
    g = np.random.default_rng(1000 + n)
    Phi, wmax = _block_diagonal_phi(g, n, max_block)
    c0, r0 = g.standard_normal(n), np.abs(g.standard_normal(n))
    MU = g.standard_normal((n, m)) * 0.1 if m else None
    cU = g.standard_normal(m) if m else None
    rU = np.abs(g.standard_normal(m)) if m else None
    rpsi = np.abs(g.standard_normal(n)) * 0.01 if m else None
    rows = np.sort(g.choice(n, size=max(3, n // 5), replace=False)).astype(np.int32) if subset else None
    Cs, Rs = ctx.bffpsv18_boxes_csc(sp.csc_matrix(Phi), c0, r0, N, MU, cU, rU, rpsi, rows)
    info = ctx.csc_info()
    assert info["sparse_path"] and 1 <= info["max_component"] <= wmax
    P = orc.DiscreteBoxProblem(Phi, c0, r0, MU, cU, rU, rpsi, None, None)
    Co, Ro = orc.bffpsv18_reach(P, N, rows)
    assert_close(Cs.T, Co, "centres (block-diagonal path)")
    assert_close(Rs.T, Ro, "radii (block-diagonal path)")
    Cd, Rd = ctx.bffpsv18_boxes(Phi, c0, r0, N, MU, cU, rU, rpsi, rows)
    assert_close(Cs, Cd, "centres vs the dense device path")
    assert_close(Rs, Rd, "radii vs the dense device path")
    assert np.all(Rs >= 0)


def test_block_diagonal_path_needs_small_components(ctx):
    """One component of 17 states is beyond the block-diagonal path: the sparse twin falls back to the dense recurrence."""
    import scipy.sparse as sp
    g = np.random.default_rng(17)
    n = 40
    Phi, _ = _block_diagonal_phi(g, n, 3, permute=False)
    Phi[:17, :17] += np.diag(np.full(16, 0.01), 1)              # chain over the first 17 states
    c0, r0 = g.standard_normal(n), np.abs(g.standard_normal(n))
    Cs, Rs = ctx.bffpsv18_boxes_csc(sp.csc_matrix(Phi), c0, r0, 25)
    assert not ctx.csc_info()["sparse_path"]
    Cd, Rd = ctx.bffpsv18_boxes(Phi, c0, r0, 25)
    assert np.array_equal(Cs, Cd) and np.array_equal(Rs, Rd)


def test_iss_shaped_modes_under_bffpsv18_full_length(ctx):
    """The decoupled modes of the ISS-shaped C2 system under BFFPSV18 (1-D blocks, N = 2000): Φ = e^{Aδ} is 135 components of
    two states, so the sparse twin runs the whole flowpipe in one launch; prefix against the oracle, every step against the
    dense device path."""
    import scipy.sparse as sp
    from reachability_jl_b200 import workloads as wl
    w = wl.iss()
    D = orc.discretize_box(w.A, orc.Box(w.c, w.r), w.delta, w.B, orc.Box(w.cU, w.rU))
    N = 2000
    Phi = np.where(np.abs(D.Phi) > 0, D.Phi, 0.0)
    Cs, Rs = ctx.bffpsv18_boxes_csc(sp.csc_matrix(Phi), D.c0, D.r0, N, D.MU, D.cU, D.rU, D.rpsi)
    info = ctx.csc_info()
    assert info["sparse_path"] and info["max_component"] == 2
    Co, Ro = orc.bffpsv18_reach(D, 150)
    assert_close(Cs[:, :150].T, Co, "centres")
    assert_close(Rs[:, :150].T, Ro, "radii")
    Cd, Rd = ctx.bffpsv18_boxes(Phi, D.c0, D.r0, N, D.MU, D.cU, D.rU, D.rpsi)
    assert_close(Cs, Cd, "centres vs the dense device path, all 2000 steps")
    assert_close(Rs, Rd, "radii vs the dense device path, all 2000 steps")
