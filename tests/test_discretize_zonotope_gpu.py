"""GPU parity: `discretize(...; set_operations="zonotope")` on the device (reach_discretize_zonotope; discretize.jl:705-743,
882-890) -- what GLGM06 feeds on (GLGM06/post.jl:15-18).

Three independent checks of the device kernels:
 (1) entry by entry (1e-10 rel + 1e-12 abs, same column order, same generator counts) against the CPU oracle, whose
     implementation shares no code with the kernels (numpy set arithmetic vs warp-per-column CUDA);
 (2) support functions: the device Ω0 must CONTAIN Z1 = X0 and Z2 = ΦX0 ⊕ δU0 ⊕ Eψ0 ⊕ Einit, whose support functions are
     evaluated lazily from the box discretisation (ρ(d, M·Box) = dᵀM c + |dᵀM| r) -- no generator formula involved -- and
     must not exceed the formula-independent bound  max(d·c1, d·c2) + (ρ(d,Z1) - d·c1) + (ρ(d,Z2) - d·c2);
 (3) assuming LazySets pairs the generators of Z1 and Z2 by index (recalled, unverifiable without Julia):
     ρ(d, Ω0) = max(d·c1, d·c2) + Σ_j max(|d·g1j|, |d·g2j|) + Σ_{j>q} |d·g1j|  -- the closed form that
     |a+b|/2 + |a-b|/2 = max(|a|, |b|) gives the hull formula.
"""
import numpy as np
import pytest

from oracle import reach_oracle as orc
from reachability_jl_b200 import ReachError
from tests.util import assert_close

pytestmark = pytest.mark.gpu


def _problem(seed, n, m, flat=False, identity_B=False):
    g = np.random.default_rng(seed)
    A = g.standard_normal((n, n)) / np.sqrt(n) - 1.2 * np.eye(n)
    c = g.standard_normal(n)
    r = np.abs(g.standard_normal(n)) * 0.1
    if flat:
        r[::3] = 0.0                          # flat directions of X0: no generator (convert(Zonotope, ::Hyperrectangle))
    if m == 0:
        return A, c, r, None, None, None
    if identity_B:
        m = n
        B = None
    else:
        B = g.standard_normal((n, m))
        if flat and m > 1:
            B[:, 1] = 0.0                     # an input that does not enter: its generator is all-zero
    cU = g.standard_normal(m)
    rU = np.abs(g.standard_normal(m)) * 0.3
    if flat:
        rU[0] = 0.0
    return A, c, r, B, cU, rU


CASES = [(4, 0, False, False), (4, 2, False, False), (9, 3, True, False), (16, 0, True, False), (33, 5, True, False),
         (12, 12, False, True), (64, 1, False, False), (130, 3, True, False)]


@pytest.mark.parametrize("rzg", [True, False])
@pytest.mark.parametrize("alg", ["forward", "backward"])
@pytest.mark.parametrize("n,m,flat,identity_B", CASES)
def test_device_zonotope_discretisation_equals_oracle(ctx, n, m, flat, identity_B, alg, rzg):
    A, c, r, B, cU, rU = _problem(40 + n + m, n, m, flat, identity_B)
    D = ctx.discretize_zonotope(A, 0.01, c, r, B, cU, rU, algorithm=alg, remove_zero_generators=rzg)
    U = None if cU is None else orc.Box(cU, rU)
    Do = orc.discretize_zonotope(A, orc.Box(c, r), 0.01, B, U, algorithm=alg, rzg=rzg)
    assert_close(D["Phi"], Do.Phi, "Phi")
    assert D["G0"].shape == Do.Omega0.G.shape, (D["G0"].shape, Do.Omega0.G.shape)
    assert_close(D["c0"], Do.Omega0.c, "centre of Omega0")
    assert_close(D["G0"], Do.Omega0.G, "generators of Omega0")
    if U is None:
        assert D["cV"] is None and D["GV"] is None
    else:
        assert D["GV"].shape == Do.V.G.shape
        assert_close(D["cV"], Do.V.c, "centre of V")
        assert_close(D["GV"], Do.V.G, "generators of V")


def _rho_zono(d, c, G):
    return d @ c + np.abs(d @ G).sum()


@pytest.mark.parametrize("n,m", [(6, 0), (6, 2), (25, 3), (40, 0)])
def test_device_hull_zonotope_contains_both_operands_and_is_tight(ctx, n, m):
    A, c, r, B, cU, rU = _problem(700 + n, n, m, flat=True)
    delta = 0.02
    D = ctx.discretize_zonotope(A, delta, c, r, B, cU, rU)
    Db = ctx.discretize_box(A, delta, c, r, B, cU, rU)                   # Φ, rE, rψ: the lazy ingredients of Z2
    Phi, rE = Db["Phi"], Db["rE"]
    g = np.random.default_rng(n)
    dirs = [g.standard_normal(n) for _ in range(200)] + [s * e for e in np.eye(n) for s in (1.0, -1.0)]
    for d in dirs:
        rho1 = d @ c + np.abs(d) @ r                                     # ρ(d, X0)
        if m:
            c2 = Phi @ c + delta * (B @ cU)
            rho2 = d @ c2 + np.abs(d @ Phi) @ r + delta * (np.abs(d @ B) @ rU) + np.abs(d) @ Db["rpsi"] + np.abs(d) @ rE
        else:
            c2 = Phi @ c
            rho2 = d @ c2 + np.abs(d @ Phi) @ (r + rE)                    # Φ also maps Einit (discretize.jl:708)
        rho0 = _rho_zono(d, D["c0"], D["G0"])
        scale = abs(rho0) + abs(rho1) + abs(rho2)
        assert rho0 >= max(rho1, rho2) - 1e-10 * scale, "Omega0 does not contain CH(Z1, Z2)"
        bound = max(d @ c, d @ c2) + (rho1 - d @ c) + (rho2 - d @ c2)
        assert rho0 <= bound + 1e-10 * scale, "Omega0 is looser than any pairing of the generators allows"


@pytest.mark.parametrize("n,m", [(5, 0), (8, 2), (21, 4)])
def test_device_hull_zonotope_closed_form_assuming_lazysets_pairs_generators_by_index(ctx, n, m):
    A, c, r, B, cU, rU = _problem(800 + n, n, m, flat=True)
    delta = 0.02
    D = ctx.discretize_zonotope(A, delta, c, r, B, cU, rU)
    Db = ctx.discretize_box(A, delta, c, r, B, cU, rU)
    Phi, rE = Db["Phi"], Db["rE"]
    keep = r != 0.0
    G1 = np.diag(r)[:, keep]
    if m:
        dU = delta * (B * rU[None, :])
        G2 = np.hstack([(Phi * r[None, :])[:, keep], dU[:, np.any(dU != 0, axis=0)], np.diag(Db["rpsi"]), np.diag(rE)])
        c2 = Phi @ c + delta * (B @ cU)
    else:
        G2 = np.hstack([(Phi * r[None, :])[:, keep], Phi * rE[None, :]])
        c2 = Phi @ c
    q = G1.shape[1]
    assert G2.shape[1] >= q
    g = np.random.default_rng(n)
    for _ in range(100):
        d = g.standard_normal(n)
        a, b = np.abs(d @ G2), np.abs(d @ G1)
        closed = max(d @ c, d @ c2) + np.maximum(a[:q], b).sum() + a[q:].sum()
        rho0 = _rho_zono(d, D["c0"], D["G0"])
        assert abs(rho0 - closed) <= 1e-10 * abs(closed) + 1e-12


def test_zonotope_discretisation_rejects_models_without_zonotope_operations(ctx):
    """discretize.jl:105-115: firstorder / nobloating with set_operations="zonotope" throw ArgumentError."""
    A, c, r, B, cU, rU = _problem(1, 5, 2)
    for alg in ("firstorder", "nobloating"):
        with pytest.raises(ValueError, match="not implemented"):
            ctx.discretize_zonotope(A, 0.01, c, r, B, cU, rU, algorithm=alg)
    with pytest.raises(ValueError, match="unknown"):
        ctx.discretize_zonotope(A, 0.01, c, r, B, cU, rU, algorithm="interval_matrix")
    # and at the ABI itself (what a ccall caller sees): status 1 with the reference's message
    import ctypes as C
    from reachability_jl_b200._lib import dptr, fmat, fvec
    Af, cf, rf = fmat(A), fvec(c), fvec(r)
    G0 = np.zeros((5, 21), order="F"); c0 = np.zeros(5)
    p0, pV = C.c_int64(), C.c_int64()
    rc = ctx.lib.reach_discretize_zonotope(ctx._h, 5, dptr(Af), 5, 0.01, 3, dptr(cf), dptr(rf), 0, None, 5, None, None, 0,
                                           None, 5, dptr(c0), dptr(G0), 5, C.byref(p0), None, None, 5, C.byref(pV))
    assert rc == 1 and b"nobloating with set operations=zonotope is not implemented" in ctx.lib.reach_last_error(ctx._h)
