"""CPU oracle for the linear-flowpipe hot path of JuliaReach/Reachability.jl v0.7.0.

TEST INFRASTRUCTURE ONLY.  Nothing under ``reachability.jl_b200/`` may import this module; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (the ``cpu_baseline`` leg and ``--impl reference``)
use it, and there only as the checker / the timed CPU arm.  The shipped product path is CUDA-only.

PARITY STATUS: **parity unpinned by the reference**.  The reference is pure Julia; no Julia toolchain exists
in this container or on the GPU box (probed: ``julia: command not found``), and the reference's own tests
hold no golden vector for Phi, Omega0, box bounds, zonotopes or Girard selections.  The
oracle is pinned instead by (tests/test_oracle_pins.py): mpmath 50-digit expm / Phi2 series, scipy expm,
closed forms, trajectory soundness, the BFFPSV18<->GLGM06 centre identity, the literal lazy-set restatement in
``oracle/lazysets_mini.py`` and the three check-mode known answers of
``/root/reference/test/Reachability/solve_continuous.jl:49-75``.

Each function cites the reference file:line (relative to /root/reference) it restates.  The set arithmetic
of the path lives in third-party dependencies that are NOT under /root/reference:
  * LazySets.jl      (Project.toml:33, compat 1.30 ... 1.36; no Manifest => exact version unpinned)
  * Julia stdlib LinearAlgebra ``exp`` (julia ~1.0/1.1/1.2, Project.toml:28)
Their published algorithms are restated here (Girard 2005 order reduction; Higham 2005 scaling-and-squaring
Pade), anchored on the reference's call sites.

All arrays are float64; matrices are handled as plain numpy 2-D arrays (layout irrelevant to the oracle).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import scipy.linalg as sla

__all__ = [
    "expm_julia", "exp_Adelta", "Pmatrix", "Phi1", "Phi2", "phi12_series",
    "Box", "Zono", "discretize_box", "discretize_zonotope", "DiscreteBoxProblem", "DiscreteZonoProblem",
    "compute_blocks", "compute_dimensions", "bffpsv18_reach", "bffpsv18_check", "bffpsv18_output_map",
    "reduce_order", "girard_selection", "glgm06_reach", "glgm06_reach_homog", "glgm06_reach_inhomog",
    "n_steps_bffpsv18", "n_steps_glgm06", "canon_zono", "ch_zonotope", "box_to_zono",
]

# ---------------------------------------------------------------------------------------------------------
# Matrix exponential: Julia stdlib `exp!(A::StridedMatrix{<:BlasFloat})` (Higham 2005), the method behind
# `exp(Matrix(A*δ))` at src/ReachSets/discretize.jl:152 and `exp(Matrix(Pmatrix(...)))` at :224, :299.
# ---------------------------------------------------------------------------------------------------------

_PADE = {
    3: [120., 60., 12., 1.],
    5: [30240., 15120., 3360., 420., 30., 1.],
    7: [17297280., 8648640., 1995840., 277200., 25200., 1512., 56., 1.],
    9: [17643225600., 8821612800., 2075673600., 302702400., 30270240., 2162160., 110880., 3960., 90., 1.],
    13: [64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800., 129060195264000.,
         10559470521600., 670442572800., 33522128640., 1323241920., 40840800., 960960., 16380., 182., 1.],
}


def expm_julia(A: np.ndarray) -> np.ndarray:
    """Restatement of Julia 1.0-1.2 ``LinearAlgebra.exp!`` for a dense real matrix.

    Hermitian input -> eigendecomposition path.  Otherwise LAPACK ``gebal('B')`` balancing, 1-norm, Pade
    degree 3/5/7/9 for norm <= 0.015/0.25/0.95/2.1, else scaling by 2^-ceil(log2(norm/5.4)) + Pade-13 +
    repeated squaring; balancing undone at the end.  
    """
    A = np.array(A, dtype=np.float64, copy=True)
    n = A.shape[0]
    assert A.shape == (n, n)
    if n == 0:
        return A
    if np.array_equal(A, A.T):
        w, V = np.linalg.eigh(A)
        return (V * np.exp(w)) @ V.T
    Ab, T = sla.matrix_balance(A, permute=True, scale=True, separate=False)
    nA = np.linalg.norm(Ab, 1)
    I = np.eye(n)
    if nA <= 2.1:
        if nA > 0.95:
            C = _PADE[9]
        elif nA > 0.25:
            C = _PADE[7]
        elif nA > 0.015:
            C = _PADE[5]
        else:
            C = _PADE[3]
        A2 = Ab @ Ab
        P = I.copy()
        U = C[1] * P
        V = C[0] * P
        for k in range(1, len(C) // 2):
            P = P @ A2
            U = U + C[2 * k + 1] * P
            V = V + C[2 * k] * P
        U = Ab @ U
        X = np.linalg.solve(V - U, V + U)
    else:
        s = math.log2(nA / 5.4)
        si = 0
        if s > 0:
            si = int(math.ceil(s))
            Ab = Ab / (2.0 ** si)
        CC = _PADE[13]
        A2 = Ab @ Ab
        A4 = A2 @ A2
        A6 = A2 @ A4
        U = Ab @ (A6 @ (CC[13] * A6 + CC[11] * A4 + CC[9] * A2) + CC[7] * A6 + CC[5] * A4 + CC[3] * A2 + CC[1] * I)
        V = A6 @ (CC[12] * A6 + CC[10] * A4 + CC[8] * A2) + CC[6] * A6 + CC[4] * A4 + CC[2] * A2 + CC[0] * I
        X = np.linalg.solve(V - U, V + U)
        for _ in range(si):
            X = X @ X
    # undo balancing: A = T * Ab * T^-1  =>  exp(A) = T * exp(Ab) * T^-1 ; T is a scaled permutation
    # (scipy returns T with A_balanced = T^-1 A T).
    Tinv = np.zeros_like(T)
    nz = T != 0
    Tinv.T[nz] = 1.0 / T[nz]
    return T @ X @ Tinv


def exp_Adelta(A: np.ndarray, delta: float) -> np.ndarray:
    """``exp_Aδ(A, δ; exp_method="base")`` = ``exp(Matrix(A*δ))`` -- src/ReachSets/discretize.jl:150-152."""
    return expm_julia(np.asarray(A, dtype=np.float64) * delta)


def Pmatrix(A: np.ndarray, delta: float) -> np.ndarray:
    """3n x 3n block matrix ``[Aδ δI 0; 0 0 δI; 0 0 0]`` -- src/ReachSets/discretize.jl:162-166."""
    n = A.shape[0]
    P = np.zeros((3 * n, 3 * n))
    P[:n, :n] = A * delta
    P[:n, n:2 * n] = delta * np.eye(n)
    P[n:2 * n, 2 * n:] = delta * np.eye(n)
    return P


def Phi1(A: np.ndarray, delta: float) -> np.ndarray:
    """``Φ₁(A, δ)`` = ``exp(Pmatrix)[1:n, n+1:2n]`` -- src/ReachSets/discretize.jl:221-226."""
    n = A.shape[0]
    return expm_julia(Pmatrix(np.asarray(A, dtype=np.float64), delta))[:n, n:2 * n]


def Phi2(A: np.ndarray, delta: float) -> np.ndarray:
    """``Φ₂(A, δ)`` = ``exp(Pmatrix)[1:n, 2n+1:3n]`` -- src/ReachSets/discretize.jl:296-301."""
    n = A.shape[0]
    return expm_julia(Pmatrix(np.asarray(A, dtype=np.float64), delta))[:n, 2 * n:]


def phi12_series(A: np.ndarray, delta: float, terms: int = 60):
    """Independent cross-check: the defining series Φ₁ = Σ δ^{i+1}/(i+1)! A^i, Φ₂ = Σ δ^{i+2}/(i+2)! A^i
    (docstrings at src/ReachSets/discretize.jl:168-176 and :243-251), summed directly with scaling/doubling
    (Φ₀(2t)=Φ₀², Φ₁(2t)=(Φ₀+I)Φ₁, Φ₂(2t)=(Φ₀+I)Φ₂+tΦ₁, read off the square of exp(Pmatrix))."""
    A = np.asarray(A, dtype=np.float64)
    n = A.shape[0]
    nrm = np.linalg.norm(A, 1) * delta
    s = max(0, int(math.ceil(math.log2(max(nrm, 1e-300) / 0.5)))) if nrm > 0.5 else 0
    h = delta / (2 ** s)
    X = A * h
    S2 = np.eye(n) / math.factorial(terms + 2)
    for j in range(terms - 1, -1, -1):  # Horner on Σ X^j/(j+2)!
        S2 = np.eye(n) / math.factorial(j + 2) + X @ S2
    S1 = np.eye(n) + X @ S2
    S0 = np.eye(n) + X @ S1
    P0, P1, P2, t = S0, h * S1, h * h * S2, h
    for _ in range(s):
        P2 = (P0 + np.eye(n)) @ P2 + t * P1
        P1 = (P0 + np.eye(n)) @ P1
        P0 = P0 @ P0
        t *= 2
    return P0, P1, P2


# ---------------------------------------------------------------------------------------------------------
# Sets on the path: boxes (c, r) and zonotopes (c, G)
# ---------------------------------------------------------------------------------------------------------

@dataclass
class Box:
    """``Hyperrectangle(c, r)`` / ``BallInf`` / product of ``Interval``s (LazySets)."""
    c: np.ndarray
    r: np.ndarray

    @property
    def lo(self):
        return self.c - self.r

    @property
    def hi(self):
        return self.c + self.r


@dataclass
class Zono:
    """``Zonotope(center, generators)`` (LazySets)."""
    c: np.ndarray
    G: np.ndarray  # n x p

    @property
    def ngens(self):
        return self.G.shape[1]


def _strip_zero_cols(G: np.ndarray) -> np.ndarray:
    """``Zonotope(c, G; remove_zero_generators=true)`` deletes all-zero columns (recalled)."""
    if G.shape[1] == 0:
        return G
    keep = np.any(G != 0.0, axis=0)
    return G[:, keep]


def make_zono(c, G, remove_zero_generators: bool = True) -> Zono:
    G = np.asarray(G, dtype=np.float64).reshape(len(c), -1)
    if remove_zero_generators:
        G = _strip_zero_cols(G)
    return Zono(np.asarray(c, dtype=np.float64), G)


def canon_zono(Z: Zono) -> Zono:
    """Canonical form used by parity comparisons: zero columns dropped."""
    return Zono(Z.c, _strip_zero_cols(Z.G))


def box_to_zono(b: Box, rzg: bool = True) -> Zono:
    """``convert(Zonotope, ::AbstractHyperrectangle)`` via `_convert_or_overapproximate`
    (src/ReachSets/discretize.jl:882-890): one axis generator per coordinate."""
    return make_zono(b.c, np.diag(b.r), rzg)


def linear_map(M: np.ndarray, Z: Zono, rzg: bool = True) -> Zono:
    """``linear_map(M, Z) = Zonotope(M*c, M*G)`` (LazySets; call sites GLGM06/reach.jl:16,48,54)."""
    return make_zono(M @ Z.c, M @ Z.G, rzg)


def minkowski_sum(Z1: Zono, Z2: Zono, rzg: bool = True) -> Zono:
    """``minkowski_sum(Z1, Z2) = Zonotope(c1+c2, hcat(G1, G2))`` (LazySets; GLGM06/reach.jl:48,54)."""
    return make_zono(Z1.c + Z2.c, np.hstack([Z1.G, Z2.G]), rzg)


def ch_zonotope(Z1: Zono, Z2: Zono, rzg: bool = True) -> Zono:
    """``overapproximate(ConvexHull(Z1, Z2), Zonotope)`` (LazySets; call sites discretize.jl:709,724;
    SURVEY.md B.5): operand with more generators first; c=(c1+c2)/2;
    G=[(G1[:,1:q]+G2)/2, (c1-c2)/2, (G1[:,1:q]-G2)/2, G1[:,q+1:end]]."""
    if Z2.ngens > Z1.ngens:
        Z1, Z2 = Z2, Z1
    c1, c2, G1, G2 = Z1.c, Z2.c, Z1.G, Z2.G
    q = G2.shape[1]
    c = (c1 + c2) / 2
    if G1.shape[1] == q:
        G = np.hstack([G1 + G2, (c1 - c2)[:, None], G1 - G2]) / 2
    else:
        G = np.hstack([(G1[:, :q] + G2) / 2, ((c1 - c2) / 2)[:, None], (G1[:, :q] - G2) / 2, G1[:, q:]])
    return make_zono(c, G, rzg)


# ---------------------------------------------------------------------------------------------------------
# discretize: forward / backward interpolation  (src/ReachSets/discretize.jl:627-743)
# ---------------------------------------------------------------------------------------------------------

@dataclass
class DiscreteBoxProblem:
    """Array form of ``IVP(CLCDS(ϕ, I, X, Ud), Ω0)`` with lazy sets (discretize.jl:665,674,678-702), already
    box-decomposed as BFFPSV18 does at reach.jl:64-65."""
    Phi: np.ndarray
    c0: np.ndarray          # decomposed Omega0 box centre (all n coordinates)
    r0: np.ndarray          # decomposed Omega0 box radius
    MU: Optional[np.ndarray]  # n x m = δ·B          (None: homogeneous)
    cU: Optional[np.ndarray]
    rU: Optional[np.ndarray]
    rpsi: Optional[np.ndarray]  # radius of Eψ0 (centre 0)
    rE: np.ndarray           # radius of Einit (centre 0)
    Phi2abs: np.ndarray


@dataclass
class DiscreteZonoProblem:
    """Array form of the ``set_operations="zonotope"`` discretisation (discretize.jl:705-743)."""
    Phi: np.ndarray
    Omega0: Zono
    V: Optional[Zono]        # discretised constant input (None: homogeneous)


def _bloating(A, Phi, X0: Box, delta, algorithm, B=None, U: Optional[Box] = None, Phi2abs=None):
    """``Einit`` / ``Eψ0`` radii: discretize.jl:652-657, :671 with `sih` = symmetric_interval_hull
    (centre 0, radius |c|+r) and box approximation of a linear map of a box (B.6)."""
    if Phi2abs is None:
        Phi2abs = Phi2(np.abs(A), delta)                      # :652
    if algorithm == "forward":
        M = A @ A                                             # (A * A) * X0, :655
    elif algorithm == "backward":
        M = A @ A @ Phi                                       # (A * A * ϕ) * X0, :657
    else:
        raise ValueError(f"the algorithm {algorithm} is unknown")   # :659
    s1 = np.abs(M @ X0.c) + np.abs(M) @ X0.r                  # inner sih
    rE = Phi2abs @ s1                                         # outer sih: |0| + |Φ₂|·s1, Φ₂(|A|) >= 0
    rpsi = None
    if U is not None:
        AB = A @ B if B is not None else A
        s2 = np.abs(AB @ U.c) + np.abs(AB) @ U.r              # sih(A * U0), :671
        rpsi = Phi2abs @ s2
    return Phi2abs, rE, rpsi


def discretize_varying_inputs(A, delta: float, B, Us: Sequence[Box], Phi2abs=None):
    """``VaryingInput`` branch of the interpolation models (discretize.jl:690-696 lazy, :729-737 zonotope): for every
    input set ``Uk`` of the sequence, ``Eψk = sih(Phi2Aabs * sih(A * Uk))`` and ``Ud[k] = δ*Uk ⊕ Eψk``.  Returns the
    list of radii of Eψk (centre 0); δB is shared by all k."""
    A = np.asarray(A, dtype=np.float64)
    if Phi2abs is None:
        Phi2abs = Phi2(np.abs(A), delta)
    AB = A @ B if B is not None else A
    out = []
    for Uk in Us:
        s = np.abs(AB @ Uk.c) + np.abs(AB) @ Uk.r             # sih(A * Uk)
        out.append(Phi2abs @ s)                               # sih(Phi2Aabs * ...): Φ₂(|A|) >= 0
    return out


def discretize_box(A, X0: Box, delta: float, B=None, U: Optional[Box] = None,
                   algorithm: str = "forward", Phi=None, Phi2abs=None) -> DiscreteBoxProblem:
    """``discretize(IVP, δ; algorithm="forward"|"backward", set_operations="lazy")`` followed by the box
    decomposition of the lazy ``Ω0 = CH(X0, ϕX0 ⊕ δU0 ⊕ Eψ0 ⊕ Einit)`` (discretize.jl:679,685) that
    ``decompose(problem.x0, partition, block_options_init)`` performs at BFFPSV18/reach.jl:64-65 for
    ``Interval``/``Hyperrectangle`` block options (per-coordinate support functions).
    ``U`` is the box the input ranges over, mapped through ``B`` (normalize: ConstantInput(B*U),
    Utils/normalization.jl:93-102,152)."""
    A = np.asarray(A, dtype=np.float64)
    n = A.shape[0]
    if Phi is None:
        Phi = exp_Adelta(A, delta)                            # :649
    if algorithm in ("firstorder", "nobloating"):
        return _discretize_box_other(A, X0, delta, B, U, algorithm, Phi)
    Phi2abs, rE, rpsi = _bloating(A, Phi, X0, delta, algorithm, B, U, Phi2abs)
    absPhi = np.abs(Phi)
    cY = Phi @ X0.c
    rY = absPhi @ X0.r + rE
    MU = cU = rU = None
    if U is not None:
        Bm = np.eye(n) if B is None else np.asarray(B, dtype=np.float64)
        MU = delta * Bm
        cU, rU = U.c, U.r
        cY = cY + MU @ cU
        rY = rY + np.abs(MU) @ rU + rpsi
    hi = np.maximum(X0.c + X0.r, cY + rY)                     # ρ(e_l, CH(X,Y)) = max
    lo = np.minimum(X0.c - X0.r, cY - rY)
    c0 = (hi + lo) / 2
    r0 = (hi - lo) / 2
    return DiscreteBoxProblem(Phi, c0, r0, MU, cU, rU, rpsi, rE, Phi2abs)


def _discretize_box_other(A, X0: Box, delta, B, U, algorithm, Phi) -> DiscreteBoxProblem:
    """The two other approximation models `discretize` dispatches to (discretize.jl:105-115), in the same array form.

    * ``"nobloating"`` (discretize.jl:527-552): ``Ω0 = X0``, ``V = Φ₁(A, δ) U`` -- no bloating terms.
    * ``"firstorder"`` with ``p = Inf`` (discretize.jl:403-473, Le Guernic & Girard 2010): ``κ = e^{δ‖A‖} − 1 − δ‖A‖``,
      ``α = κ (‖X0‖ + ‖U‖/‖A‖)``, ``β = κ ‖U‖/‖A‖``, ``Ω0 = CH(X0, ΦX0 ⊕ δU ⊕ □α)``, ``V = δU ⊕ □β`` with infinity-norm
      balls, i.e. uniform box radii (``norm(X, Inf)`` of a box = max_i |c_i| + r_i; of the lazy ``B*U`` its box bound)."""
    n = A.shape[0]
    Bm = None if U is None else (np.eye(n) if B is None else np.asarray(B, dtype=np.float64))
    if algorithm == "nobloating":
        if U is None:
            return DiscreteBoxProblem(Phi, X0.c.copy(), X0.r.copy(), None, None, None, None, np.zeros(n), None)
        MU = Phi1(A, delta) @ Bm                                 # Φ₁(A, δ) * (B U)
        return DiscreteBoxProblem(Phi, X0.c.copy(), X0.r.copy(), MU, U.c, U.r, np.zeros(n), np.zeros(n), None)
    Anorm = np.abs(A).sum(axis=1).max()                       # opnorm(A, Inf)
    kappa = math.exp(delta * Anorm) - 1.0 - delta * Anorm
    RX0 = np.max(np.abs(X0.c) + X0.r)
    cY = Phi @ X0.c
    rY = np.abs(Phi) @ X0.r
    MU = cU = rU = rpsi = None
    if U is None:
        alpha = kappa * RX0
    else:
        RU = np.max(np.abs(Bm @ U.c) + np.abs(Bm) @ U.r)
        alpha = kappa * (RX0 + RU / Anorm)
        beta = kappa * RU / Anorm
        MU, cU, rU = delta * Bm, U.c, U.r
        rpsi = np.full(n, beta)
        cY = cY + MU @ cU
        rY = rY + np.abs(MU) @ rU
    rE = np.full(n, alpha)
    rY = rY + rE
    hi = np.maximum(X0.c + X0.r, cY + rY)
    lo = np.minimum(X0.c - X0.r, cY - rY)
    return DiscreteBoxProblem(Phi, (hi + lo) / 2, (hi - lo) / 2, MU, cU, rU, rpsi, rE, None)


def discretize_zonotope(A, X0: Box, delta: float, B=None, U: Optional[Box] = None,
                        algorithm: str = "forward", rzg: bool = True, Phi=None, Phi2abs=None) -> DiscreteZonoProblem:
    """``discretize(...; set_operations="zonotope")`` -- discretize.jl:705-743 (GLGM06/post.jl:15-18)."""
    A = np.asarray(A, dtype=np.float64)
    n = A.shape[0]
    if Phi is None:
        Phi = exp_Adelta(A, delta)
    Phi2abs, rE, rpsi = _bloating(A, Phi, X0, delta, algorithm, B, U, Phi2abs)
    Einit = box_to_zono(Box(np.zeros(n), rE), rzg)            # :706 / :715
    Z1 = box_to_zono(X0, rzg)                                  # :707 / :717
    if U is None:
        Z2 = linear_map(Phi, minkowski_sum(Z1, Einit, rzg), rzg)   # :708 (Φ also maps Einit)
        return DiscreteZonoProblem(Phi, ch_zonotope(Z1, Z2, rzg), None)   # :709
    Bm = np.eye(n) if B is None else np.asarray(B, dtype=np.float64)
    Epsi0 = box_to_zono(Box(np.zeros(n), rpsi), rzg)          # :716
    PhiX0 = linear_map(Phi, Z1, rzg)                           # :718
    U0 = make_zono(Bm @ U.c, Bm @ np.diag(U.r), rzg)           # :719  convert(Zonotope, B*U)
    dU0 = linear_map(delta * np.eye(n), U0, rzg)               # :720-721
    Z2 = minkowski_sum(minkowski_sum(minkowski_sum(PhiX0, dU0, rzg), Epsi0, rzg), Einit, rzg)   # :722
    Omega0 = ch_zonotope(Z1, Z2, rzg)                          # :723
    V = minkowski_sum(dU0, Epsi0, rzg)                         # :726
    return DiscreteZonoProblem(Phi, Omega0, V)


# ---------------------------------------------------------------------------------------------------------
# BFFPSV18  (src/ReachSets/ContinuousPost/BFFPSV18/*)
# ---------------------------------------------------------------------------------------------------------

def n_steps_bffpsv18(T: float, delta: float) -> int:
    """``N = ceil(Int, T / δ)`` -- BFFPSV18/reach.jl:54."""
    return int(math.ceil(T / delta))


def n_steps_glgm06(T: float, delta: float) -> int:
    """``N = round(Int, T / δ)`` -- GLGM06/post.jl:12 (Julia rounds half to even, like Python)."""
    return int(round(T / delta))


def compute_blocks(vars_1b: Sequence[int], partition: Sequence[Sequence[int]]):
    """``compute_blocks(vars, partition)`` -- BFFPSV18/BFFPSV18.jl:413-434.  1-based in, 1-based out."""
    blocks = []
    nxt = 0
    var_idx = 0
    vars_1b = list(vars_1b)
    for i, block in enumerate(partition, start=1):
        nxt += len(block)
        if vars_1b[var_idx] <= nxt:
            blocks.append(i)
            var_idx += 1
            while var_idx < len(vars_1b) and vars_1b[var_idx] <= nxt:
                var_idx += 1
            if var_idx >= len(vars_1b):
                break
    assert var_idx == len(vars_1b)
    return blocks


def compute_dimensions(partition, blocks):
    """``compute_dimensions(partition, blocks)`` -- BFFPSV18/compute_dimensions.jl:1-15 (1-based)."""
    dims = []
    for b in blocks:
        dims.extend(list(partition[b - 1]))
    return dims


def bffpsv18_reach(P: DiscreteBoxProblem, N: int, rows: Optional[Sequence[int]] = None):
    """Dense ``reach_blocks!`` (BFFPSV18/reach_blocks.jl:135-224) in array form, for
    ``Interval``/``Hyperrectangle`` block options, ``ConstantInput`` or no input, ``Universe`` invariant
    (termination_N, reach.jl:225-227).  ``rows``: 0-based coordinates of the blocks of interest.

    Returns (centers, radii), each (N, len(rows)): step k (1-based) in row k-1.
    """
    Phi = P.Phi
    n = Phi.shape[0]
    rows = np.arange(n) if rows is None else np.asarray(rows, dtype=np.int64)
    C = np.zeros((N, len(rows)))
    R = np.zeros((N, len(rows)))
    C[0] = P.c0[rows]                                         # :152-158
    R[0] = P.r0[rows]
    if N == 1:
        return C, R                                            # reach.jl:87-90
    inhomog = P.MU is not None
    if inhomog:
        cW = (P.MU @ P.cU)[rows]                               # :172-175 box of proj(bi)*inputs
        rW = (np.abs(P.MU) @ P.rU + P.rpsi)[rows]
    Pk = Phi[rows, :].copy()                                   # ϕpowerk = copy(ϕ), only needed rows
    for k in range(2, N + 1):
        c = Pk @ P.c0                                          # :185-193
        r = np.abs(Pk) @ P.r0
        if inhomog:
            c = c + cW
            r = r + rW
        C[k - 1] = c                                           # :195-202
        R[k - 1] = r
        if k == N:                                             # :204-207
            break
        if inhomog:                                            # :209-215
            PM = Pk @ P.MU
            cW = cW + PM @ P.cU
            rW = rW + np.abs(PM) @ P.rU + np.abs(Pk) @ P.rpsi
        Pk = Pk @ Phi                                          # :217-218
    return C, R


def bffpsv18_reach_lazy(A, P: DiscreteBoxProblem, delta: float, N: int, rows: Optional[Sequence[int]] = None):
    """``reach_blocks!(ϕ::SparseMatrixExp, ...)`` (BFFPSV18/reach_blocks.jl:227-310), restated LITERALLY: the lazy
    exponent matrix accumulates, ``ϕpowerk.M .= ϕpowerk.M + ϕ.M`` (:306), and the rows of ``exp(ϕpowerk.M)`` are taken
    from scratch at every step (``row(ϕpowerk, bi) = get_rows(ϕpowerk, bi)``, :37-38 -- LazySets evaluates them with
    Expokit's Krylov ``expmv``; here with a dense ``expm`` of kδA, i.e. to full precision).  Same box arithmetic as
    `bffpsv18_reach`; used to pin the device path's recurrence W_k = W_{k-1} e^{δA} against the from-scratch rows.
    O(N n^3): small cases only."""
    A = np.asarray(A.toarray() if hasattr(A, "toarray") else A, dtype=np.float64)
    n = A.shape[0]
    rows = np.arange(n) if rows is None else np.asarray(rows, dtype=np.int64)
    C = np.zeros((N, len(rows)))
    R = np.zeros((N, len(rows)))
    C[0] = P.c0[rows]
    R[0] = P.r0[rows]
    if N == 1:
        return C, R
    inhomog = P.MU is not None
    if inhomog:
        cW = (P.MU @ P.cU)[rows]                               # :256-262
        rW = (np.abs(P.MU) @ P.rU + P.rpsi)[rows]
    M1 = A * delta                                             # ϕ.M
    Mk = M1.copy()                                             # ϕpowerk = SparseMatrixExp(copy(ϕ.M))  :257
    for k in range(2, N + 1):
        Pk = expm_julia(Mk)[rows, :]                           # row(ϕpowerk, bi)  :269
        c = Pk @ P.c0
        r = np.abs(Pk) @ P.r0
        if inhomog:
            c = c + cW
            r = r + rW
        C[k - 1] = c
        R[k - 1] = r
        if inhomog:                                            # :284-287 (inside the block loop, before the store)
            PM = Pk @ P.MU
            cW = cW + PM @ P.cU
            rW = rW + np.abs(PM) @ P.rU + np.abs(Pk) @ P.rpsi
        Mk = Mk + M1                                           # :306
    return C, R


def _block_slices(nrows: int, block_sizes: Optional[Sequence[int]]):
    """Slices of ``rows`` belonging to each block of ``:blocks`` (consecutive groups; None: one single block)."""
    if block_sizes is None:
        return [slice(0, nrows)]
    assert sum(block_sizes) == nrows
    out, o = [], 0
    for b in block_sizes:
        out.append(slice(o, o + b))
        o += b
    return out


def _blockwise_abs_product(M: np.ndarray, Pk: np.ndarray, slices) -> np.ndarray:
    """Σ_i |M[:, b_i] · P[b_i, :]|: the support function of a ``CartesianProductArray`` of lazy
    ``MinkowskiSumArray`` blocks is separable per block, ρ(d, ×_i X_i) = Σ_i ρ(d_i, X_i), so the absolute value
    is taken per block of rows (NOT of the full product M·P)."""
    acc = np.zeros((M.shape[0], Pk.shape[1]))
    for sl in slices:
        acc += np.abs(M[:, sl] @ Pk[sl, :])
    return acc


def bffpsv18_output_map(P: DiscreteBoxProblem, N: int, M: np.ndarray, rows: Sequence[int],
                        block_sizes: Optional[Sequence[int]] = None):
    """``output_function`` branch of ``reach_blocks!`` (reach_blocks.jl:152-155,191-199; reach.jl:73-82):
    blocks stay lazy (``MinkowskiSumArray``) and the stored set is ``box_approximation(M * CPA(Xhatk))``:
    centre ``M(P c0 + cW)``, radius ``Σ_i |M[:,b_i] P[b_i,:]| r0 + |M| rW`` (per-block absolute value, see
    `_blockwise_abs_product`; step 1 uses the boxed ``Xhat0``: ``|M| r0[rows]``, :152-155).
    ``M`` is q x len(rows); ``block_sizes`` groups ``rows`` into the blocks of ``:blocks``."""
    Phi = P.Phi
    rows = np.asarray(rows, dtype=np.int64)
    slices = _block_slices(len(rows), block_sizes)
    q = M.shape[0]
    C = np.zeros((N, q))
    R = np.zeros((N, q))
    C[0] = M @ P.c0[rows]
    R[0] = np.abs(M) @ P.r0[rows]
    if N == 1:
        return C, R
    inhomog = P.MU is not None
    if inhomog:
        cW = (P.MU @ P.cU)[rows]
        rW = (np.abs(P.MU) @ P.rU + P.rpsi)[rows]
    Pk = Phi[rows, :].copy()
    for k in range(2, N + 1):
        c = (M @ Pk) @ P.c0
        r = _blockwise_abs_product(M, Pk, slices) @ P.r0
        if inhomog:
            c = c + M @ cW
            r = r + np.abs(M) @ rW
        C[k - 1], R[k - 1] = c, r
        if k == N:
            break
        if inhomog:
            PM = Pk @ P.MU
            cW = cW + PM @ P.cU
            rW = rW + np.abs(PM) @ P.rU + np.abs(Pk) @ P.rpsi
        Pk = Pk @ Phi
    return C, R


def bffpsv18_check(P: DiscreteBoxProblem, N: int, ell: np.ndarray, bound: float, rows: Sequence[int],
                   eager: bool = True, block_sizes: Optional[Sequence[int]] = None) -> int:
    """Dense ``check_blocks`` (BFFPSV18/check_blocks.jl:105-183) for the half-space property
    ``is_contained_in(LinearConstraint(ell, bound))`` after `inout_map_property` (partitions.jl:33-46):
    states stay lazy (no per-block boxing) and ρ of the ``CartesianProductArray`` is separable per block, so
    ``ρ(ℓ, X_k) = (ℓᵀP)c0 + Σ_i |ℓ_iᵀP[b_i,:]| r0 + ℓ·cW + |ℓ|·rW`` (SURVEY.md section 8f-1 misses the
    per-block absolute value; `oracle/lazysets_mini.check_blocks_dense` is the literal version).  ``ell`` is
    given on the coordinates ``rows`` (0-based); ``block_sizes`` groups ``rows`` into the blocks of ``:blocks``.
    Returns the first violating step index (1-based) or 0 if the property holds for all N steps."""
    rows = np.asarray(rows, dtype=np.int64)
    ell = np.asarray(ell, dtype=np.float64)
    slices = _block_slices(len(rows), block_sizes)
    violation = 0
    if ell @ P.c0[rows] + np.abs(ell) @ P.r0[rows] > bound:   # check_blocks.jl:122-129, k = 1
        if eager:
            return 1
        violation = 1
    if N == 1:
        return violation
    inhomog = P.MU is not None
    if inhomog:
        cW = (P.MU @ P.cU)[rows]
        rW = (np.abs(P.MU) @ P.rU + P.rpsi)[rows]
    Pk = P.Phi[rows, :].copy()
    for k in range(2, N + 1):
        rho = (ell @ Pk) @ P.c0 + _blockwise_abs_product(ell[None, :], Pk, slices)[0] @ P.r0
        if inhomog:
            rho += ell @ cW + np.abs(ell) @ rW
        if rho > bound:
            if eager:
                return k
            if violation == 0:
                violation = k
        if k == N:
            break
        if inhomog:
            PM = Pk @ P.MU
            cW = cW + PM @ P.cU
            rW = rW + np.abs(PM) @ P.rU + np.abs(Pk) @ P.rpsi
        Pk = Pk @ P.Phi
    return violation


# ---------------------------------------------------------------------------------------------------------
# GLGM06  (src/ReachSets/ContinuousPost/GLGM06/*)  +  LazySets.reduce_order (Girard 2005)
# ---------------------------------------------------------------------------------------------------------

def girard_selection(G: np.ndarray, r):
    """Scores and permutation of ``reduce_order`` (LazySets): ``h_j = ‖g_j‖₁ − ‖g_j‖∞``,
    ``ind = sortperm(h)`` (stable: ties by index), ``m = p − floor(n (r−1))`` discarded.
    Returns (h, ind (0-based), m)."""
    n, p = G.shape
    aG = np.abs(G)
    h = aG.sum(axis=0) - aG.max(axis=0)
    ind = np.argsort(h, kind="stable")
    m = p - int(math.floor(n * (r - 1)))
    return h, ind, m


def reduce_order(Z: Zono, r, rzg: bool = True, return_selection: bool = False):
    """``reduce_order(Z::Zonotope, r)`` -- LazySets (call sites GLGM06/post.jl:20, reach.jl:49,55).
    ``r*n >= p`` or ``r < 1``: unchanged; else discard the m lowest-score generators into a box
    ``Diagonal(Σ|g|)`` appended AFTER the kept generators (kept in ascending-score order)."""
    n, p = Z.G.shape
    if r * n >= p or r < 1:
        return (Z, None) if return_selection else Z
    h, ind, m = girard_selection(Z.G, r)
    d = np.abs(Z.G[:, ind[:m]]).sum(axis=1)
    if m < p:
        Gred = np.hstack([Z.G[:, ind[m:]], np.diag(d)])
    else:
        Gred = np.diag(d)
    out = make_zono(Z.c, Gred, rzg)
    return (out, (h, ind, m)) if return_selection else out


def glgm06_reach_homog(Omega0: Zono, Phi: np.ndarray, N: int, rzg: bool = True):
    """``reach_homog!`` -- GLGM06/reach.jl:5-24: ``R_k = linear_map(Φ, R_{k-1})``."""
    R = [Omega0]
    for _ in range(2, N + 1):
        R.append(linear_map(Phi, R[-1], rzg))
    return R


def glgm06_iter_inhomog(Omega0: Zono, V: Zono, Phi: np.ndarray, N: int, max_order: int, rzg: bool = True):
    """``reach_inhomog!`` -- GLGM06/reach.jl:30-62, one reach set at a time: yields ``(k, R_k, (selR, selW))`` for
    k = 1..N (selections None at k = 1), so that full-length parity tests need not hold the whole flowpipe."""
    yield 1, Omega0, None
    W = V                                                     # :41-42
    Pk = Phi.copy()                                           # :43
    for k in range(2, N + 1):
        Rk = minkowski_sum(linear_map(Pk, Omega0, rzg), W, rzg)          # :48
        Rk, selR = reduce_order(Rk, max_order, rzg, return_selection=True)   # :49
        W = minkowski_sum(W, linear_map(Pk, V, rzg), rzg)               # :54
        W, selW = reduce_order(W, max_order, rzg, return_selection=True)   # :55
        yield k, Rk, (selR, selW)
        Pk = Pk @ Phi                                         # :57-58


def glgm06_reach_inhomog(Omega0: Zono, V: Zono, Phi: np.ndarray, N: int, max_order: int, rzg: bool = True,
                         return_selections: bool = False):
    """``reach_inhomog!`` -- GLGM06/reach.jl:30-62."""
    R, sels = [], []
    for _, Rk, sel in glgm06_iter_inhomog(Omega0, V, Phi, N, max_order, rzg):
        R.append(Rk)
        sels.append(sel)
    return (R, sels) if return_selections else R


def glgm06_reach(D: DiscreteZonoProblem, N: int, max_order: int = 10, rzg: bool = True,
                 return_selections: bool = False):
    """``post(::GLGM06, ...)`` after discretisation -- GLGM06/post.jl:20-37."""
    Omega0red = reduce_order(D.Omega0, max_order, rzg)        # :20
    if D.V is None:
        R = glgm06_reach_homog(Omega0red, D.Phi, N, rzg)      # :31-32
        return (R, [None] * N) if return_selections else R
    return glgm06_reach_inhomog(Omega0red, D.V, D.Phi, N, max_order, rzg, return_selections)   # :34-36
