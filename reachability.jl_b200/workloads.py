"""Synthetic benchmark systems of BASELINE.json's five configs.

Deterministic (`numpy.random.Generator(PCG64(seed))`); the same arrays feed the oracle, the CUDA path and the
bench, so all three see identical bits.  Shapes follow the named models (building, ISS, FOM, MNA5, heli28);
the real model files live in external repositories that are not under /root/reference.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np


@dataclass
class Workload:
    name: str
    A: np.ndarray                      # n x n (dense float64)
    c: np.ndarray                      # X0 box centre
    r: np.ndarray                      # X0 box radius
    B: Optional[np.ndarray]            # n x m or None (homogeneous)
    cU: Optional[np.ndarray]
    rU: Optional[np.ndarray]
    delta: float
    T: float
    algorithm: str                     # "BFFPSV18" | "GLGM06"
    partition: Optional[list] = None   # BFFPSV18: list of 0-based coordinate lists
    max_order: int = 10
    extra: dict = field(default_factory=dict)

    @property
    def n(self):
        return self.A.shape[0]

    @property
    def m(self):
        return 0 if self.B is None else self.B.shape[1]


def _rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def building(n_dof: int = 24, delta: float = 0.003, T: float = 20.0, seed: int = 48) -> Workload:
    """C1: building-shaped second-order chain, n = 2*n_dof (48), one input, 1-D Interval blocks."""
    g = _rng(seed)
    k = g.uniform(50.0, 500.0, n_dof + 1)
    K = np.diag(k[:-1] + k[1:]) - np.diag(k[1:-1], 1) - np.diag(k[1:-1], -1)
    D = 0.02 * K + 0.1 * np.eye(n_dof)
    n = 2 * n_dof
    A = np.zeros((n, n))
    A[:n_dof, n_dof:] = np.eye(n_dof)
    A[n_dof:, :n_dof] = -K
    A[n_dof:, n_dof:] = -D
    B = np.zeros((n, 1))
    B[n_dof, 0] = 1.0
    c = np.zeros(n)
    r = np.zeros(n)
    r[:min(10, n)] = 2.5e-4
    r[n_dof] = 1e-4
    return Workload("C1-building", A, c, r, B, np.array([0.9]), np.array([0.1]), delta, T, "BFFPSV18",
                    partition=[[i] for i in range(n)])


def iss(n_modes: int = 135, delta: float = 0.01, T: float = 20.0, seed: int = 270, max_order: int = 10) -> Workload:
    """C2: ISS-shaped lightly damped modal system, n = 2*n_modes (270), three inputs, GLGM06."""
    g = _rng(seed)
    w = np.exp(g.uniform(np.log(0.8), np.log(60.0), n_modes))
    zeta = 0.005
    n = 2 * n_modes
    A = np.zeros((n, n))
    A[:n_modes, n_modes:] = np.eye(n_modes)
    A[n_modes:, :n_modes] = -np.diag(w ** 2)
    A[n_modes:, n_modes:] = -np.diag(2 * zeta * w)
    B = np.zeros((n, 3))
    B[n_modes:, :] = g.normal(0.0, 1e-3, (n_modes, 3))
    c = np.zeros(n)
    r = np.full(n, 1e-4)
    cU = np.array([0.05, 0.9, 0.95])
    rU = np.array([0.05, 0.1, 0.05])
    return Workload("C2-iss", A, c, r, B, cU, rU, delta, T, "GLGM06", max_order=max_order)


def iss_mixed(n_modes: int = 512, delta: float = 0.01, T: float = 6.0, seed: int = 1024, max_order: int = 10) -> Workload:
    """A large n * p GLGM06 case (what the generator-column sharding exists for): the lightly damped
    modes of `iss` under a random orthogonal change of coordinates, so A -- and every power of Phi -- is DENSE and the
    generator map is the DMMA GEMM.  n = 2 * n_modes (1024), three inputs, same reachable set as the decoupled modes up to
    the rotation."""
    g = _rng(seed)
    w = np.exp(g.uniform(np.log(0.8), np.log(60.0), n_modes))
    zeta = 0.005
    n = 2 * n_modes
    Am = np.zeros((n, n))
    Am[:n_modes, n_modes:] = np.eye(n_modes)
    Am[n_modes:, :n_modes] = -np.diag(w ** 2)
    Am[n_modes:, n_modes:] = -np.diag(2 * zeta * w)
    Q, _ = np.linalg.qr(g.standard_normal((n, n)))
    A = Q @ Am @ Q.T
    B = Q @ np.vstack([np.zeros((n_modes, 3)), g.normal(0.0, 1e-3, (n_modes, 3))])
    c = np.zeros(n)
    r = np.full(n, 1e-4)
    cU = np.array([0.05, 0.9, 0.95])
    rU = np.array([0.05, 0.1, 0.05])
    return Workload(f"GLGM06-dense-n{n}", A, c, r, B, cU, rU, delta, T, "GLGM06", max_order=max_order)


def fom(n_pairs: int = 500, delta: float = 0.003, T: float = 20.0, seed: int = 1006) -> Workload:
    """C3: FOM-shaped sparse stable system, n = 6 + 2*n_pairs (1006), one input, 2-D Hyperrectangle blocks."""
    g = _rng(seed)
    n = 6 + 2 * n_pairs
    A = np.zeros((n, n))
    Q, _ = np.linalg.qr(g.standard_normal((6, 6)))
    A[:6, :6] = Q @ (np.diag(-g.uniform(1.0, 10.0, 6)) + np.triu(g.standard_normal((6, 6)), 1)) @ Q.T
    a = g.uniform(1.0, 100.0, n_pairs)
    w = g.uniform(1.0, 300.0, n_pairs)
    for i in range(n_pairs):
        o = 6 + 2 * i
        A[o, o] = -a[i]
        A[o, o + 1] = w[i]
        A[o + 1, o] = -w[i]
        A[o + 1, o + 1] = -a[i]
    # 0.1 % random off-diagonal coupling, scaled so every Gershgorin disc stays in Re < 0
    nnz_c = max(1, int(1e-3 * n * n))
    ii = g.integers(0, n, nnz_c)
    jj = g.integers(0, n, nnz_c)
    vals = g.standard_normal(nnz_c)
    Cm = np.zeros((n, n))
    for i_, j_, v_ in zip(ii, jj, vals):
        if i_ != j_ and A[i_, j_] == 0.0:
            Cm[i_, j_] = v_
    margin = -np.diag(A)
    rowsum = np.abs(Cm).sum(axis=1)
    scale = np.where(rowsum > 0, 0.1 * margin / np.maximum(rowsum, 1e-300), 0.0)
    A += Cm * scale[:, None]
    B = g.standard_normal((n, 1))
    c = np.zeros(n)
    r = np.zeros(n)
    r[:min(400, n)] = 1e-4
    c[min(400, n):min(800, n)] = 2.5e-4
    part = [[2 * i, 2 * i + 1] for i in range(n // 2)]
    return Workload("C3-fom", A, c, r, B, np.array([0.0]), np.array([1.0]), delta, T, "BFFPSV18", partition=part)


def mna5(n: int = 10913, delta: float = 1e-3, N: int = 2000, seed: int = 10913) -> Workload:
    """C4: MNA5-shaped sparse system A = -(diag + graph Laplacian, ~5 nnz/row), nine unit inputs, 1-D blocks."""
    g = _rng(seed)
    A = np.zeros((n, n))
    d = g.uniform(1.0, 1e3, n)
    for _ in range(2):      # two random symmetric neighbours per row (+ diagonal) ~ 5 nnz/row
        j = g.integers(0, n, n)
        wgt = g.uniform(0.1, 10.0, n)
        for i in range(n):
            if i != j[i]:
                A[i, j[i]] += wgt[i]
                A[j[i], i] += wgt[i]
    L = np.diag(A.sum(axis=1)) - A
    A = -(np.diag(d) + L)
    m = 9
    B = np.zeros((n, m))
    for q in range(m):
        B[q * (n // m), q] = 1.0
    c = np.zeros(n)
    r = np.zeros(n)
    c[:10] = 2e-4
    r[:10] = 1e-4
    base = np.linspace(0.1, 0.9, m)
    return Workload("C4-mna5", A, c, r, B, base, 0.1 * base, delta, N * delta, "BFFPSV18",
                    partition=[[i] for i in range(n)], extra={"N": N})


def heli28(n: int = 28, delta: float = 0.01, T: float = 20.0, seed: int = 28) -> Workload:
    """C5 base system: heli28-shaped dense stable homogeneous system (ensemble splits via `split_box`)."""
    g = _rng(seed)
    assert n % 2 == 0
    re = g.uniform(-5.0, -0.05, n // 2)
    im = g.uniform(0.0, 20.0, n // 2)
    Tq = np.zeros((n, n))
    for i in range(n // 2):
        Tq[2 * i, 2 * i] = re[i]
        Tq[2 * i, 2 * i + 1] = im[i]
        Tq[2 * i + 1, 2 * i] = -im[i]
        Tq[2 * i + 1, 2 * i + 1] = re[i]
    Tq += np.triu(g.normal(0.0, 0.5, (n, n)), 2)
    Q, _ = np.linalg.qr(g.standard_normal((n, n)))
    A = Q @ Tq @ Q.T
    c = np.zeros(n)
    r = np.zeros(n)
    r[:8] = 0.1
    return Workload("C5-heli28", A, c, r, None, None, None, delta, T, "GLGM06")


def split_box(c, r, splits):
    """Split box (c, r) into prod(splits) sub-boxes; splits[i] = number of equal parts along coordinate i."""
    c = np.asarray(c, float)
    r = np.asarray(r, float)
    n = len(c)
    splits = list(splits) + [1] * (n - len(splits))
    grids = []
    for i in range(n):
        k = splits[i]
        w = r[i] / k
        grids.append(c[i] - r[i] + w * (2 * np.arange(k) + 1))
    mesh = np.stack(np.meshgrid(*grids, indexing="ij"), axis=0).reshape(n, -1)
    rr = np.tile((r / np.array(splits))[:, None], (1, mesh.shape[1]))
    return mesh, rr


CONFIGS = {"C1": building, "C2": iss, "C3": fom, "C4": mna5, "C5": heli28}
