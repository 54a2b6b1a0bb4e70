"""Literal (object-graph) restatement of the reference's lazy-set hot path, for SMALL cases only.

TEST INFRASTRUCTURE ONLY (see oracle/reach_oracle.py header).  Pure-Python loops; use n <= ~12.

Where ``reach_oracle.py`` states the path on arrays, this file restates it the way the
reference executes it: tiny lazy-set classes with support functions ``ρ`` and the same double loop over
blocks as ``reach_blocks!`` (src/ReachSets/ContinuousPost/BFFPSV18/reach_blocks.jl:135-224), the same lazy
``Ω0`` expression as ``_discretize_interpolation_{homog,inhomog}(…, Val(:lazy))``
(src/ReachSets/discretize.jl:678-702) and the same ``decompose`` call (BFFPSV18/reach.jl:64-65).  Agreement
of the two is one of the oracle's pins: it checks the array-form derivation, not LazySets itself (LazySets.jl is
not under /root/reference; its ρ rules are restated from SURVEY.md B.6-B.8).
"""
from __future__ import annotations

import numpy as np


class LazySet:
    def rho(self, d):  # support function ρ(d, X)
        raise NotImplementedError

    def dim(self):
        raise NotImplementedError


class Hyperrectangle(LazySet):
    def __init__(self, c, r):
        self.c = np.asarray(c, dtype=np.float64)
        self.r = np.asarray(r, dtype=np.float64)

    def rho(self, d):
        return float(d @ self.c + np.abs(d) @ self.r)

    def dim(self):
        return len(self.c)


class Interval(Hyperrectangle):
    def __init__(self, lo, hi):
        super().__init__([(lo + hi) / 2], [(hi - lo) / 2])
        self.lo, self.hi = lo, hi


class ZeroSet(Hyperrectangle):
    def __init__(self, n):
        super().__init__(np.zeros(n), np.zeros(n))


class LinearMap(LazySet):
    def __init__(self, M, X):
        self.M = np.atleast_2d(np.asarray(M, dtype=np.float64))
        self.X = X

    def rho(self, d):
        return self.X.rho(self.M.T @ d)

    def dim(self):
        return self.M.shape[0]


class MinkowskiSumArray(LazySet):
    def __init__(self, arr):
        self.arr = list(arr)

    def rho(self, d):
        return sum(X.rho(d) for X in self.arr)

    def dim(self):
        return self.arr[0].dim()


class ConvexHull(LazySet):
    def __init__(self, X, Y):
        self.X, self.Y = X, Y

    def rho(self, d):
        return max(self.X.rho(d), self.Y.rho(d))

    def dim(self):
        return self.X.dim()


class CartesianProductArray(LazySet):
    def __init__(self, arr):
        self.arr = list(arr)

    def rho(self, d):
        s, o = 0.0, 0
        for X in self.arr:
            k = X.dim()
            s += X.rho(d[o:o + k])
            o += k
        return s

    def dim(self):
        return sum(X.dim() for X in self.arr)


def box_approximation(X: LazySet) -> Hyperrectangle:
    """``overapproximate(X, Hyperrectangle)``."""
    n = X.dim()
    c, r = np.zeros(n), np.zeros(n)
    for i in range(n):
        e = np.zeros(n)
        e[i] = 1.0
        htop = X.rho(e)
        hbottom = -X.rho(-e)
        c[i] = (htop + hbottom) / 2
        r[i] = (htop - hbottom) / 2
    return Hyperrectangle(c, r)


def overapproximate_interval(X: LazySet) -> Interval:
    """``overapproximate(X, Interval)`` for a 1-D set."""
    assert X.dim() == 1
    return Interval(-X.rho(np.array([-1.0])), X.rho(np.array([1.0])))


def symmetric_interval_hull(X: LazySet) -> Hyperrectangle:
    """``symmetric_interval_hull(X)``: centre 0, radius |c|+r of the box approximation."""
    b = box_approximation(X)
    return Hyperrectangle(np.zeros(b.dim()), np.abs(b.c) + b.r)


def overapprox(X: LazySet):
    return overapproximate_interval(X) if X.dim() == 1 else box_approximation(X)


def decompose(X: LazySet, partition):
    """``decompose(X, partition, Interval|Hyperrectangle)``; partition is 0-based lists."""
    n = X.dim()
    out = []
    for blk in partition:
        Pm = np.zeros((len(blk), n))
        for a, l in enumerate(blk):
            Pm[a, l] = 1.0
        out.append(overapprox(LinearMap(Pm, X)))
    return out


def discretize_lazy(A, X0: Hyperrectangle, delta, Phi, Phi2abs, B=None, U: Hyperrectangle | None = None,
                    algorithm="forward"):
    """discretize.jl:627-702 with lazy sets.  Returns (Omega0 lazy set, V lazy set or None)."""
    A = np.asarray(A, dtype=np.float64)
    n = A.shape[0]
    sih = symmetric_interval_hull
    if algorithm == "forward":
        Einit = sih(LinearMap(Phi2abs, sih(LinearMap(A @ A, X0))))          # :655
    else:
        Einit = sih(LinearMap(Phi2abs, sih(LinearMap(A @ A @ Phi, X0))))    # :657
    if U is None:
        return ConvexHull(X0, MinkowskiSumArray([LinearMap(Phi, X0), Einit])), None      # :679
    Bm = np.eye(n) if B is None else np.asarray(B, dtype=np.float64)
    U0 = LinearMap(Bm, U)                                                    # normalization.jl:152
    Epsi0 = sih(LinearMap(Phi2abs, sih(LinearMap(A, U0))))                   # :671
    dU0 = LinearMap(delta * np.eye(n), U0)
    Omega0 = ConvexHull(X0, MinkowskiSumArray([LinearMap(Phi, X0), dU0, Epsi0, Einit]))   # :685
    V = MinkowskiSumArray([dU0, Epsi0])                                      # :688
    return Omega0, V


def reach_blocks_dense(Phi, Xhat0, V, N, blocks, partition):
    """``reach_blocks!`` dense method, reach_blocks.jl:135-224, ``output_function == nothing``,
    ``overapproximate`` = Interval for 1-D blocks / Hyperrectangle otherwise, lazy_inputs_interval=always
    (reach.jl:115-116).  ``blocks`` 0-based block indices; ``partition`` 0-based coordinate lists.
    Returns list over steps of lists of block sets."""
    n = Phi.shape[0]
    res = [[Xhat0[b] for b in blocks]]                                       # :152-158
    if N == 1:
        return res
    Pk = Phi.copy()
    if V is not None:
        What = []
        for b in blocks:                                                     # :169-176
            bi = partition[b]
            Pm = np.zeros((len(bi), n))
            for a, l in enumerate(bi):
                Pm[a, l] = 1.0
            What.append(overapprox(LinearMap(Pm, V)))
    k = 2
    while True:
        Xhatk = []
        for i, b in enumerate(blocks):                                       # :184-194
            bi = partition[b]
            arr = [LinearMap(Pk[np.ix_(bi, bj)], Xhat0[j]) for j, bj in enumerate(partition)]
            if V is not None:
                arr.append(What[i])
            Xhatk.append(overapprox(MinkowskiSumArray(arr)))
        res.append(Xhatk)                                                    # :195-202
        if k >= N:                                                           # :204-207
            break
        if V is not None:                                                    # :209-215
            for i, b in enumerate(blocks):
                bi = partition[b]
                What[i] = overapprox(MinkowskiSumArray([What[i], LinearMap(Pk[bi, :], V)]))
        Pk = Pk @ Phi                                                        # :217-218
        k += 1
    return res


def check_blocks_dense(Phi, Xhat0, V, N, blocks, partition, ell, bound, eager=True):
    """``check_blocks`` dense, check_blocks.jl:105-183, property ρ(ℓ, X) <= bound on the coordinates of
    ``blocks``.  Returns first violating step (1-based) or 0."""
    n = Phi.shape[0]
    ell = np.asarray(ell, dtype=np.float64)
    violation = 0
    if CartesianProductArray([Xhat0[b] for b in blocks]).rho(ell) > bound:
        if eager:
            return 1
        violation = 1
    if N == 1:
        return violation
    Pk = Phi.copy()
    if V is not None:
        What = []
        for b in blocks:
            bi = partition[b]
            Pm = np.zeros((len(bi), n))
            for a, l in enumerate(bi):
                Pm[a, l] = 1.0
            What.append(overapprox(LinearMap(Pm, V)))
    k = 2
    while True:
        Xhatk = []
        for i, b in enumerate(blocks):
            bi = partition[b]
            arr = [LinearMap(Pk[np.ix_(bi, bj)], Xhat0[j]) for j, bj in enumerate(partition)]
            if V is not None:
                arr.append(What[i])
            Xhatk.append(MinkowskiSumArray(arr))
        if CartesianProductArray(Xhatk).rho(ell) > bound:
            if eager:
                return k
            if violation == 0:
                violation = k
        if k == N:
            break
        if V is not None:
            for i, b in enumerate(blocks):
                bi = partition[b]
                What[i] = overapprox(MinkowskiSumArray([What[i], LinearMap(Pk[bi, :], V)]))
        Pk = Pk @ Phi
        k += 1
    return violation
