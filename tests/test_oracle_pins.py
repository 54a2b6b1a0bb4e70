"""Pins of the CPU oracle (oracle/reach_oracle.py).

The reference is pure Julia and cannot run here or on the GPU box, and its own tests hold no numeric golden
vector for this path => "parity unpinned by the reference".  The oracle is pinned by
  (1) 50-digit mpmath expm / series, (2) scipy expm, (3) closed forms, (4) the literal lazy-set restatement
  (oracle/lazysets_mini.py), (5) trajectory soundness, (6) BFFPSV18 <-> GLGM06 identities,
  (7) the three check-mode known answers of /root/reference/test/Reachability/solve_continuous.jl:49-75.
"""
import math

import mpmath as mp
import numpy as np
import pytest
import scipy.linalg as sla

from oracle import lazysets_mini as lz
from oracle import reach_oracle as orc
from tests.util import assert_close, tol_units

A4 = np.array([[0.0509836, 0.168159, 0.95246, 0.33644],
               [0.42377, 0.67972, 0.129232, 0.126662],
               [0.518654, 0.981313, 0.489854, 0.588326],
               [0.38318, 0.616014, 0.518412, 0.778765]])   # test/Reachability/solve_continuous.jl:6-9


def _mp_expm(A, digits=50):
    mp.mp.dps = digits
    return np.array(mp.expm(mp.matrix(A.tolist())).tolist(), dtype=np.float64)


def _mp_phi12(A, delta, digits=50, terms=80):
    mp.mp.dps = digits
    n = A.shape[0]
    Am = mp.matrix(A.tolist())
    P1 = mp.zeros(n)
    P2 = mp.zeros(n)
    Ai = mp.eye(n)
    for i in range(terms):
        P1 += Ai * (mp.mpf(delta) ** (i + 1) / mp.factorial(i + 1))
        P2 += Ai * (mp.mpf(delta) ** (i + 2) / mp.factorial(i + 2))
        Ai = Ai * Am
    f = lambda M: np.array(M.tolist(), dtype=np.float64)
    return f(P1), f(P2)


@pytest.mark.parametrize("scale", [0.005, 0.1, 0.5, 1.5, 4.0, 40.0])
def test_expm_julia_vs_mpmath_and_scipy(scale):
    """Every Pade branch (3/5/7/9/13 + squaring) of the Julia `exp!` restatement."""
    g = np.random.default_rng(7)
    A = g.standard_normal((7, 7))
    A *= scale / np.linalg.norm(A, 1)
    E = orc.expm_julia(A)
    ref = _mp_expm(A)
    assert np.max(np.abs(E - ref)) <= 1e-13 * max(1.0, np.max(np.abs(ref)))
    assert np.max(np.abs(E - sla.expm(A))) <= 1e-12 * max(1.0, np.max(np.abs(ref)))


def test_expm_julia_symmetric_branch():
    g = np.random.default_rng(8)
    A = g.standard_normal((6, 6))
    A = -(A @ A.T)
    assert np.max(np.abs(orc.expm_julia(A * 0.1) - _mp_expm(A * 0.1))) <= 1e-13


def test_expm_badly_scaled_needs_balancing():
    g = np.random.default_rng(9)
    A = g.standard_normal((5, 5))
    D = np.diag([1e-4, 1.0, 1e3, 1e-2, 10.0])
    A = D @ A @ np.linalg.inv(D) * 0.3
    E, ref = orc.expm_julia(A), _mp_expm(A)
    assert np.max(np.abs(E - ref) / (np.abs(ref) + 1e-300)) < 1e-8   # entries span 14 decades
    assert np.max(np.abs(E - ref)) <= 1e-12 * np.max(np.abs(ref))


@pytest.mark.parametrize("delta", [0.003, 0.01, 0.25])
def test_phi1_phi2_blocks_vs_series(delta):
    """Φ₁, Φ₂ as blocks of exp(Pmatrix) (discretize.jl:221-226, 296-301) = the defining series (:168-176, :243-251)."""
    g = np.random.default_rng(11)
    A = g.standard_normal((5, 5)) * 3.0
    P1m, P2m = _mp_phi12(A, delta)
    assert_close(orc.Phi1(A, delta), P1m, "Phi1")
    assert_close(orc.Phi2(A, delta), P2m, "Phi2")
    P0s, P1s, P2s = orc.phi12_series(A, delta)
    assert_close(P1s, P1m, "Phi1 series")
    assert_close(P2s, P2m, "Phi2 series")
    assert_close(P0s, _mp_expm(A * delta), "Phi0 series")
    # call site discretize.jl:652 uses Φ₂(|A|, δ): non-negative
    assert np.all(orc.Phi2(np.abs(A), delta) >= 0)


def test_closed_form_diagonal_and_rotation():
    a, w, delta = 0.7, 3.0, 0.01
    A = np.array([[-a, w], [-w, -a]])
    Phi = orc.exp_Adelta(A, delta)
    e = math.exp(-a * delta)
    ref = e * np.array([[math.cos(w * delta), math.sin(w * delta)], [-math.sin(w * delta), math.cos(w * delta)]])
    assert_close(Phi, ref, "rotation expm")
    # homogeneous 1-D x' = -a x, X0 = [0.9, 1.1]: box k is [e^{-a(k-1)δ} * hull...]; radius decays as Φ^{k-1} r0
    A1 = np.array([[-a]])
    D = orc.discretize_box(A1, orc.Box(np.array([1.0]), np.array([0.1])), delta)
    C, R = orc.bffpsv18_reach(D, 50)
    k = np.arange(50)
    assert_close(C[:, 0], D.c0[0] * np.exp(-a * delta * k), "1-D centres")
    assert_close(R[:, 0], D.r0[0] * np.exp(-a * delta * k), "1-D radii")
    # soundness of the first box: contains x(t) for t in [0, δ], x0 in X0
    assert D.c0[0] + D.r0[0] >= 1.1 and D.c0[0] - D.r0[0] <= 0.9 * math.exp(-a * delta)


def _small_problem(seed, n=6, m=2, homog=False):
    g = np.random.default_rng(seed)
    A = g.standard_normal((n, n)) - 1.5 * np.eye(n)
    X0 = orc.Box(g.standard_normal(n), np.abs(g.standard_normal(n)) * 0.1)
    if homog:
        return A, X0, None, None
    B = g.standard_normal((n, m))
    U = orc.Box(g.standard_normal(m), np.abs(g.standard_normal(m)) * 0.2)
    return A, X0, B, U


@pytest.mark.parametrize("homog", [False, True])
@pytest.mark.parametrize("algorithm", ["forward", "backward"])
@pytest.mark.parametrize("partition", [[[0], [1], [2], [3], [4], [5]], [[0, 1], [2, 3], [4, 5]], [[0, 1, 2], [3], [4, 5]]])
def test_array_form_equals_lazy_set_restatement(homog, algorithm, partition):
    """SURVEY.md Appendix A array form == the reference's object-graph execution (lazysets_mini)."""
    A, X0, B, U = _small_problem(3, homog=homog)
    delta, N = 0.02, 12
    D = orc.discretize_box(A, X0, delta, B, U, algorithm=algorithm)
    Om, V = lz.discretize_lazy(A, lz.Hyperrectangle(X0.c, X0.r), delta, D.Phi, D.Phi2abs, B,
                               None if U is None else lz.Hyperrectangle(U.c, U.r), algorithm)
    Xhat0 = lz.decompose(Om, partition)
    c0 = np.concatenate([x.c for x in Xhat0])
    r0 = np.concatenate([x.r for x in Xhat0])
    assert_close(D.c0, c0, "c0")
    assert_close(D.r0, r0, "r0")
    blocks = [0, 2]
    rows = [l for b in blocks for l in partition[b]]
    res = lz.reach_blocks_dense(D.Phi, Xhat0, V, N, blocks, partition)
    C, R = orc.bffpsv18_reach(D, N, rows)
    for k in range(N):
        ck = np.concatenate([x.c for x in res[k]])
        rk = np.concatenate([x.r for x in res[k]])
        assert_close(C[k], ck, f"centres step {k + 1}")
        assert_close(R[k], rk, f"radii step {k + 1}")


def test_check_mode_known_answers_of_the_reference():
    """test/Reachability/solve_continuous.jl:49-75: X0 = BallInf(ones(4), 0.1), T = 0.1, default δ = 0.01,
    partition [1:2, 3:4]; x1 + x2 <= 2 is violated at index 1 (with blocks [1] and [1, 2]); x1 - x2 <= 2 holds
    (check_blocks returns 0, BFFPSV18.jl:400-402 reports it as violation == -1)."""
    X0 = orc.Box(np.ones(4), np.full(4, 0.1))
    delta = 0.01
    N = orc.n_steps_bffpsv18(0.1, delta)
    assert N == 10
    D = orc.discretize_box(A4, X0, delta)
    part1 = [[1, 2], [3, 4]]
    # :vars => [1, 2] -> block 1 only; property restricted to block 1 coordinates (partitions.jl:33-46)
    assert orc.compute_blocks([1, 2], part1) == [1]
    assert orc.bffpsv18_check(D, N, np.array([1.0, 1.0]), 2.0, rows=[0, 1]) == 1
    assert orc.compute_blocks([1, 2, 3], part1) == [1, 2]
    assert orc.bffpsv18_check(D, N, np.array([1.0, 1.0, 0.0, 0.0]), 2.0, rows=[0, 1, 2, 3], block_sizes=[2, 2]) == 1
    assert orc.bffpsv18_check(D, N, np.array([1.0, -1.0, 0.0, 0.0]), 2.0, rows=[0, 1, 2, 3], block_sizes=[2, 2]) == 0
    # and the literal object-graph version agrees
    Om, V = lz.discretize_lazy(A4, lz.Hyperrectangle(X0.c, X0.r), delta, D.Phi, D.Phi2abs)
    part0 = [[0, 1], [2, 3]]
    Xhat0 = lz.decompose(Om, part0)
    assert lz.check_blocks_dense(D.Phi, Xhat0, V, N, [0], part0, [1.0, 1.0], 2.0) == 1
    assert lz.check_blocks_dense(D.Phi, Xhat0, V, N, [0, 1], part0, [1.0, -1.0, 0.0, 0.0], 2.0) == 0


def test_check_mode_array_form_vs_lazy():
    A, X0, B, U = _small_problem(5)
    delta, N = 0.05, 25
    D = orc.discretize_box(A, X0, delta, B, U)
    part = [[0, 1], [2, 3], [4, 5]]
    Om, V = lz.discretize_lazy(A, lz.Hyperrectangle(X0.c, X0.r), delta, D.Phi, D.Phi2abs, B, lz.Hyperrectangle(U.c, U.r))
    Xhat0 = lz.decompose(Om, part)
    g = np.random.default_rng(0)
    for trial in range(6):
        ell = g.standard_normal(4)
        C, R = orc.bffpsv18_reach(D, N, [0, 1, 4, 5])
        vals = C @ ell + R @ np.abs(ell)      # boxed bound is an upper bound of the lazy ρ
        bound = float(np.quantile(vals, 0.5))
        for eager in (True, False):
            got = orc.bffpsv18_check(D, N, ell, bound, rows=[0, 1, 4, 5], eager=eager, block_sizes=[2, 2])
            want = lz.check_blocks_dense(D.Phi, Xhat0, V, N, [0, 2], part, ell, bound, eager=eager)
            assert got == want


def test_output_map_branch_vs_lazy():
    """reach_blocks.jl:191-199 with output_function = M: box_approximation(M * CPA(lazy Xhatk))."""
    A, X0, B, U = _small_problem(6)
    delta, N = 0.03, 9
    D = orc.discretize_box(A, X0, delta, B, U)
    part = [[0, 1], [2, 3], [4, 5]]
    Om, V = lz.discretize_lazy(A, lz.Hyperrectangle(X0.c, X0.r), delta, D.Phi, D.Phi2abs, B, lz.Hyperrectangle(U.c, U.r))
    Xhat0 = lz.decompose(Om, part)
    M = np.random.default_rng(1).standard_normal((2, 4))
    C, R = orc.bffpsv18_output_map(D, N, M, [0, 1, 2, 3], block_sizes=[2, 2])
    # literal: rebuild the lazy MinkowskiSumArray per block, then box_approximation(M * CPA)
    n = 6
    Pk = D.Phi.copy()
    blocks = [0, 1]
    What = []
    for b in blocks:
        Pm = np.zeros((2, n))
        for a, l in enumerate(part[b]):
            Pm[a, l] = 1.0
        What.append(lz.overapprox(lz.LinearMap(Pm, V)))
    b1 = lz.box_approximation(lz.LinearMap(M, lz.CartesianProductArray([Xhat0[b] for b in blocks])))
    assert_close(C[0], b1.c)
    assert_close(R[0], b1.r)
    for k in range(2, N + 1):
        Xk = []
        for i, b in enumerate(blocks):
            arr = [lz.LinearMap(Pk[np.ix_(part[b], bj)], Xhat0[j]) for j, bj in enumerate(part)]
            arr.append(What[i])
            Xk.append(lz.MinkowskiSumArray(arr))
        bk = lz.box_approximation(lz.LinearMap(M, lz.CartesianProductArray(Xk)))
        assert_close(C[k - 1], bk.c, f"centre {k}")
        assert_close(R[k - 1], bk.r, f"radius {k}")
        for i, b in enumerate(blocks):
            What[i] = lz.overapprox(lz.MinkowskiSumArray([What[i], lz.LinearMap(Pk[part[b], :], V)]))
        Pk = Pk @ D.Phi


def _simulate(A, B, x0, us, delta, substeps=20):
    """Exact piecewise-constant-input trajectory sampled `substeps` times per step."""
    n = A.shape[0]
    h = delta / substeps
    Ph = sla.expm(A * h)
    if B is not None:
        m = B.shape[1]
        Mx = np.zeros((n + m, n + m))
        Mx[:n, :n] = A * h
        Mx[:n, n:] = B * h
        Gh = sla.expm(Mx)[:n, n:]
    xs = [x0]
    x = x0
    for u in us:
        for _ in range(substeps):
            x = Ph @ x + (Gh @ u if B is not None else 0.0)
            xs.append(x)
    return np.array(xs)


@pytest.mark.parametrize("homog", [False, True])
def test_soundness_boxes_and_zonotopes_contain_trajectories(homog):
    A, X0, B, U = _small_problem(12, homog=homog)
    delta, N, sub = 0.02, 30, 10
    D = orc.discretize_box(A, X0, delta, B, U)
    C, R = orc.bffpsv18_reach(D, N)
    Dz = orc.discretize_zonotope(A, X0, delta, B, U)
    Z = orc.glgm06_reach(Dz, N, max_order=3)
    g = np.random.default_rng(2)
    for trial in range(20):
        x0 = X0.c + X0.r * g.choice([-1.0, 1.0], size=len(X0.c)) * g.uniform(0.5, 1.0)
        us = [] if homog else None
        if homog:
            us = [None] * N
        else:
            u = U.c + U.r * g.choice([-1.0, 1.0], size=len(U.c))
            us = [u] * N        # constant input
        xs = _simulate(A, B, x0, us, delta, sub)
        for k in range(N):
            seg = xs[k * sub:(k + 1) * sub + 1]
            assert np.all(seg <= C[k] + R[k] + 1e-12) and np.all(seg >= C[k] - R[k] - 1e-12), f"box {k + 1}"
            # zonotope membership (necessary condition via box hull and random support directions)
            Gk = Z[k].G
            for d in g.standard_normal((5, len(x0))):
                rho = d @ Z[k].c + np.abs(d @ Gk).sum()
                assert np.all(seg @ d <= rho + 1e-12), f"zonotope {k + 1}"


def test_glgm06_identities_with_bffpsv18():
    """With no order reduction the zonotope centres follow c_k = P_{k-1} c(Ω0) + Σ P_j cV (GLGM06/reach.jl:48,54)
    and the box hull of every zonotope is contained in ... the same recurrence structure as BFFPSV18 boxes."""
    A, X0, B, U = _small_problem(21)
    delta, N = 0.02, 15
    Dz = orc.discretize_zonotope(A, X0, delta, B, U)
    Z, sels = orc.glgm06_reach(Dz, N, max_order=10 ** 6, return_selections=True)
    Om, V = Dz.Omega0, Dz.V
    Pk = np.eye(len(X0.c))
    cW = np.zeros(len(X0.c))
    for k in range(N):
        if k > 0:
            Pk = Pk @ Dz.Phi if k > 1 else Dz.Phi.copy()
            cW = cW + (np.linalg.matrix_power(Dz.Phi, k - 1) @ V.c)
            assert_close(Z[k].c, Pk @ Om.c + cW, f"centre {k + 1}")
            assert Z[k].ngens == Om.ngens + k * V.ngens
    # reduction keeps the centre and over-approximates: box hull grows
    Zr = orc.glgm06_reach(Dz, N, max_order=2)
    for k in range(N):
        assert_close(Zr[k].c, Z[k].c, "centre under reduction")
        assert np.all(np.abs(Zr[k].G).sum(axis=1) >= np.abs(Z[k].G).sum(axis=1) * (1 - 1e-12))
        assert Zr[k].ngens <= 2 * len(X0.c)


def test_reduce_order_girard_semantics():
    g = np.random.default_rng(4)
    n, p, r = 5, 23, 3
    Z = orc.Zono(g.standard_normal(n), g.standard_normal((n, p)))
    out, (h, ind, m) = orc.reduce_order(Z, r, return_selection=True)
    assert m == p - n * (r - 1) and out.ngens == n * r
    assert np.all(np.diff(h[ind]) >= 0)
    # kept generators in ascending-score order, box last
    assert np.array_equal(out.G[:, :p - m], Z.G[:, ind[m:]])
    assert np.array_equal(out.G[:, p - m:], np.diag(np.abs(Z.G[:, ind[:m]]).sum(axis=1)))
    # unchanged when r*n >= p
    assert orc.reduce_order(Z, 5) is Z
    # stable ties: axis-aligned generators all have h == 0 and keep index order
    G = np.hstack([np.diag(np.arange(1.0, n + 1)), g.standard_normal((n, 12))])
    h2, ind2, _ = orc.girard_selection(G, 2)
    assert list(ind2[:n]) == list(range(n)) and np.all(h2[:n] == 0)
    # zero generators are stripped by the constructor unless asked otherwise
    Zz = orc.Zono(np.zeros(3), np.array([[1.0, 0, 0, 2, 0.5, 1, 1], [0, 0, 0, 1, 0.5, 2, 1], [0, 0, 0, 0, 0, 0, 0]]))
    o1 = orc.reduce_order(Zz, 2)
    o2 = orc.reduce_order(Zz, 2, rzg=False)
    assert o2.ngens == 6 and o1.ngens < 6


def test_step_counts():
    assert orc.n_steps_bffpsv18(20.0, 0.003) == 6667      # ceil, BFFPSV18/reach.jl:54
    assert orc.n_steps_glgm06(20.0, 0.01) == 2000          # round, GLGM06/post.jl:12
    assert orc.compute_dimensions([[1, 2], [3, 4], [5]], [1, 3]) == [1, 2, 5]
    assert orc.compute_blocks([1, 3], [[1], [2], [3], [4]]) == [1, 3]


def test_firstorder_and_nobloating_models():
    """`discretize_firstorder` (discretize.jl:403-473) must contain sampled trajectories on every [k delta, (k+1) delta];
    `discretize_nobloating` (:527-552) is exact at the grid points for constant inputs (Phi1 = int_0^delta e^{As} ds)."""
    g = np.random.default_rng(17)
    n, m, delta, N = 5, 2, 0.05, 12
    A = g.standard_normal((n, n)) - 1.5 * np.eye(n)
    B = g.standard_normal((n, m))
    X0 = orc.Box(g.standard_normal(n), np.abs(g.standard_normal(n)) * 0.2)
    U = orc.Box(g.standard_normal(m), np.abs(g.standard_normal(m)) * 0.3)
    Dn = orc.discretize_box(A, X0, delta, B, U, algorithm="nobloating")
    Cn, Rn = orc.bffpsv18_reach(Dn, N)
    Df = orc.discretize_box(A, X0, delta, B, U, algorithm="firstorder")
    Cf, Rf = orc.bffpsv18_reach(Df, N)
    import scipy.linalg as sla
    for trial in range(20):
        x0 = X0.c + X0.r * g.uniform(-1, 1, n)
        u = U.c + U.r * g.uniform(-1, 1, m)
        # exact solution with constant input: x(t) = e^{At} x0 + int_0^t e^{As} ds B u
        M = np.zeros((n + m, n + m)); M[:n, :n] = A; M[:n, n:] = B
        for k in range(N):
            for frac in (0.0, 0.3, 0.77, 1.0):
                t = (k + frac) * delta
                E = sla.expm(M * t)
                x = E[:n, :n] @ x0 + E[:n, n:] @ u
                assert np.all(np.abs(x - Cf[k]) <= Rf[k] * (1 + 1e-9) + 1e-12), ("firstorder", k, frac)
            tg = k * delta                     # nobloating: the k-th set is the exact image at t = k delta
            E = sla.expm(M * tg)
            x = E[:n, :n] @ x0 + E[:n, n:] @ u
            assert np.all(np.abs(x - Cn[k]) <= Rn[k] * (1 + 1e-9) + 1e-12), ("nobloating", k)


def test_lazy_expm_literal_restatement_agrees_with_the_recurrence():
    """reach_blocks! for a SparseMatrixExp (reach_blocks.jl:227-310) takes the rows of exp(kδA) from scratch at every
    step; the dense variant (:135-224) multiplies by Φ.  Same sets up to rounding: the oracle's two restatements agree
    far inside the parity tolerance."""
    g = np.random.default_rng(11)
    n = 14
    A = g.standard_normal((n, n)) / 3 - 1.5 * np.eye(n)
    X0 = orc.Box(g.standard_normal(n), np.abs(g.standard_normal(n)) * 0.1)
    B = g.standard_normal((n, 2))
    U = orc.Box(g.standard_normal(2), np.array([0.1, 0.3]))
    D = orc.discretize_box(A, X0, 0.03, B, U)
    rows = [0, 3, 13]
    C1, R1 = orc.bffpsv18_reach(D, 30, rows)
    C2, R2 = orc.bffpsv18_reach_lazy(A, D, 0.03, 30, rows)
    assert_close(C2, C1, "centres")
    assert_close(R2, R1, "radii")


def test_hull_zonotope_hand_derived_example():
    """`overapproximate(ConvexHull(Z1, Z2), Zonotope)` (call sites discretize.jl:709,724) on numbers small enough to do by
    hand (Girard-style enclosure): Z1 = ((0,0), [e1 e2]) has more generators than Z2 = ((2,0), [(1,1)]), so it goes
    first and q = 1:  c = (1,0);  G = [ (e1+(1,1))/2 | ((0,0)-(2,0))/2 | (e1-(1,1))/2 | e2 ].
    The COLUMN ORDER is the recalled LazySets convention -- an assumption no reference-held value can confirm here."""
    Z = orc.ch_zonotope(orc.Zono(np.zeros(2), np.eye(2)), orc.Zono(np.array([2.0, 0.0]), np.array([[1.0], [1.0]])))
    assert np.array_equal(Z.c, [1.0, 0.0])
    assert np.array_equal(Z.G, np.array([[1.0, -1.0, 0.0, 0.0], [0.5, 0.0, -0.5, 1.0]]))
    # whatever the column order: the enclosure contains both operands' vertices
    for v in ([1, 1], [1, -1], [-1, 1], [-1, -1], [3, 1], [1, -1]):
        d = np.array([0.3, -0.8])
        assert d @ Z.c + np.abs(d @ Z.G).sum() >= d @ np.array(v, dtype=float) - 1e-15


def test_reduce_order_hand_derived_example():
    """`reduce_order(Z, r)` (call sites GLGM06/post.jl:20, reach.jl:49,55) by hand, n = 2, p = 5, r = 2:
    scores h = |g|_1 - |g|_inf = [0, 0, 2, 0.5, 1]; stable sortperm = [1, 2, 4, 5, 3] (the two axis-aligned generators tie at
    exactly 0 and keep their order); m = p - floor(n (r - 1)) = 3 generators are boxed: d = |g1| + |g2| + |g4| = (1.5, 3.5);
    result [g5 g3 | diag(d)] -- kept generators in ascending-score order, box generators last."""
    G = np.array([[1.0, 0.0, 2.0, 0.5, 1.0], [0.0, 3.0, 2.0, 0.5, 4.0]])
    Z, (h, ind, m) = orc.reduce_order(orc.Zono(np.zeros(2), G), 2, return_selection=True)
    assert np.array_equal(h, [0.0, 0.0, 2.0, 0.5, 1.0]) and list(ind) == [0, 1, 3, 4, 2] and m == 3
    assert np.array_equal(Z.G, np.array([[1.0, 2.0, 1.5, 0.0], [4.0, 2.0, 0.0, 3.5]]))
    assert orc.reduce_order(orc.Zono(np.zeros(2), G[:, :4]), 2).G.shape == (2, 4)        # r n >= p: unchanged
