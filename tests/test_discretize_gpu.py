"""GPU parity: device discretisation (exp_Aδ, Φ₁, Φ₂, discretize_interpolation + box decomposition of Ω0;
src/ReachSets/discretize.jl:150-166, 221-315, 627-675 and BFFPSV18/reach.jl:64-65) vs the CPU oracle, which
restates Julia's Padé `exp` and is itself pinned against mpmath (tests/test_oracle_pins.py).

Tolerance: 1e-10 relative + 1e-12 absolute on every entry.
"""
from pathlib import Path

import numpy as np
import pytest

from oracle import reach_oracle as orc
from tests.util import assert_close

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).resolve().parent / "golden"


def _A(seed, n, scale=1.0):
    g = np.random.default_rng(seed)
    return scale * (g.standard_normal((n, n)) / np.sqrt(n) - 1.2 * np.eye(n))


@pytest.mark.parametrize("n,delta,scale", [(1, 0.01, 1.0), (4, 0.01, 1.0), (5, 0.5, 3.0), (17, 0.1, 30.0), (48, 0.003, 100.0),
                                           (130, 0.01, 10.0), (270, 0.01, 5.0)])
def test_expm_and_phi12(ctx, n, delta, scale):
    A = _A(n, n, scale)
    Phi = ctx.expm(A, delta)
    assert_close(Phi, orc.exp_Adelta(A, delta), "exp(A delta)")
    P1, P2 = ctx.phi12(A, delta)
    _, S1, S2 = orc.phi12_series(A, delta)
    assert_close(P1, S1, "Phi1 (series)")
    assert_close(P2, S2, "Phi2 (series)")
    if n <= 48:       # the reference's own definition through the dense 3n x 3n exponential
        assert_close(P1, orc.Phi1(A, delta), "Phi1 (Pmatrix)")
        assert_close(P2, orc.Phi2(A, delta), "Phi2 (Pmatrix)")


def test_expm_symmetric_matrix(ctx):
    """Julia's `exp` takes the Hermitian eigendecomposition branch for symmetric input."""
    g = np.random.default_rng(3)
    M = g.standard_normal((20, 20))
    A = -(M @ M.T) / 20
    assert_close(ctx.expm(A, 0.05), orc.exp_Adelta(A, 0.05), "exp of a symmetric matrix")


@pytest.mark.parametrize("name", ["lcs4", "clccs4", "doc5"])
@pytest.mark.parametrize("alg", ["forward", "backward"])
def test_discretize_box_golden(ctx, name, alg):
    z = np.load(GOLDEN / f"{name}.npz")
    inh = "B" in z
    out = ctx.discretize_box(z["A"], float(z["delta"]), z["c"], z["r"], z["B"] if inh else None,
                             z["cU"] if inh else None, z["rU"] if inh else None, algorithm=alg, want_phi2abs=True)
    for key in ("Phi", "Phi2abs", "c0", "r0", "rE"):
        assert_close(out[key], z[f"{alg}_{key}"], f"{name} {alg} {key}")
    if inh:
        assert_close(out["rpsi"], z[f"{alg}_rpsi"], "rpsi")
        assert_close(out["MU"], z[f"{alg}_MU"], "MU")


@pytest.mark.parametrize("n,m", [(3, 0), (7, 2), (33, 5), (130, 1), (64, 64)])
@pytest.mark.parametrize("alg", ["forward", "backward"])
def test_discretize_box_random_then_flowpipe(ctx, n, m, alg):
    g = np.random.default_rng(n * 7 + m)
    A = _A(n + 1, n, 2.0)
    X0 = orc.Box(g.standard_normal(n), np.abs(g.standard_normal(n)) * 0.1)
    B = g.standard_normal((n, m)) if m and m != n else None
    U = orc.Box(g.standard_normal(m), np.abs(g.standard_normal(m)) * 0.3) if m else None
    D = orc.discretize_box(A, X0, 0.02, B, U, algorithm=alg)
    out = ctx.discretize_box(A, 0.02, X0.c, X0.r, B, None if U is None else U.c, None if U is None else U.r,
                             algorithm=alg, want_phi2abs=True)
    assert_close(out["Phi"], D.Phi, "Phi")
    assert_close(out["Phi2abs"], D.Phi2abs, "Phi2abs")
    assert_close(out["rE"], D.rE, "rE")
    assert_close(out["c0"], D.c0, "c0")
    assert_close(out["r0"], D.r0, "r0")
    if m:
        assert_close(out["rpsi"], D.rpsi, "rpsi")
        assert_close(out["MU"], D.MU, "MU")
    # device discretisation feeding the device flowpipe == oracle end to end
    N = 40
    C, R = ctx.bffpsv18_boxes(out["Phi"], out["c0"], out["r0"], N, out["MU"], out["cU"], out["rU"], out["rpsi"])
    Co, Ro = orc.bffpsv18_reach(D, N)
    assert_close(C.T, Co, "centres")
    assert_close(R.T, Ro, "radii")


def test_unknown_algorithm_raises(ctx):
    with pytest.raises(ValueError):
        ctx.discretize_box(np.eye(2), 0.1, np.zeros(2), np.ones(2), algorithm="sideways")


@pytest.mark.parametrize("n,m", [(3, 0), (4, 2), (33, 5), (20, 20), (130, 1)])
@pytest.mark.parametrize("alg", ["firstorder", "nobloating"])
def test_other_discretisation_models_then_flowpipe(ctx, n, m, alg):
    """`discretize_firstorder` (p = Inf, discretize.jl:403-473) and `discretize_nobloating` (:527-552) in array form."""
    g = np.random.default_rng(n * 11 + m)
    A = _A(n + 2, n, 2.0)
    X0 = orc.Box(g.standard_normal(n), np.abs(g.standard_normal(n)) * 0.1)
    B = g.standard_normal((n, m)) if m and m != n else None
    U = orc.Box(g.standard_normal(m), np.abs(g.standard_normal(m)) * 0.3) if m else None
    D = orc.discretize_box(A, X0, 0.02, B, U, algorithm=alg)
    out = ctx.discretize_box(A, 0.02, X0.c, X0.r, B, None if U is None else U.c, None if U is None else U.r, algorithm=alg)
    assert_close(out["Phi"], D.Phi, "Phi")
    assert_close(out["c0"], D.c0, "c0")
    assert_close(out["r0"], D.r0, "r0")
    assert_close(out["rE"], D.rE, "rE")
    if m:
        assert_close(out["rpsi"], D.rpsi, "rpsi")
        assert_close(out["MU"], D.MU, "MU")
    N = 30
    C, R = ctx.bffpsv18_boxes(out["Phi"], out["c0"], out["r0"], N, out["MU"], out["cU"], out["rU"], out["rpsi"])
    Co, Ro = orc.bffpsv18_reach(D, N)
    assert_close(C.T, Co, "centres")
    assert_close(R.T, Ro, "radii")
    if alg == "nobloating":
        assert np.array_equal(out["c0"], X0.c) and np.array_equal(out["r0"], X0.r)


@pytest.mark.parametrize("n,m,K,identity_B", [(4, 2, 1, False), (17, 3, 9, False), (48, 1, 40, False), (12, 12, 5, True), (130, 4, 300, False)])
def test_varying_inputs_discretisation(ctx, n, m, K, identity_B):
    """`VaryingInput` (discretize.jl:690-696, 729-737): Eψk = sih(Φ₂(|A|, δ) sih(A Uk)) for a sequence of input boxes, all k at
    once on the device vs the oracle's per-k loop; k = 1 agrees with the constant-input discretisation."""
    g = np.random.default_rng(n + K)
    A = _A(n + 1, n, 2.0)
    B = None if identity_B else g.standard_normal((n, m))
    cUs = g.standard_normal((m, K))
    rUs = np.abs(g.standard_normal((m, K))) * 0.3
    rUs[0, ::3] = 0.0
    out = ctx.discretize_inputs_box(A, 0.02, B, cUs, rUs)
    want = orc.discretize_varying_inputs(A, 0.02, B, [orc.Box(cUs[:, k], rUs[:, k]) for k in range(K)])
    assert_close(out["rpsi"], np.stack(want, axis=1), "radii of E_psi_k")
    assert_close(out["MU"], 0.02 * (np.eye(n) if B is None else B), "delta * B")
    const = ctx.discretize_box(A, 0.02, np.zeros(n), np.zeros(n), B, cUs[:, 0], rUs[:, 0])
    assert_close(out["rpsi"][:, 0], const["rpsi"], "k = 1 vs the ConstantInput discretisation")
