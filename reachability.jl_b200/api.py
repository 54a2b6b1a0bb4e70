"""Host-side mirror of the reference's operator interface for the linear-flowpipe hot path.

Same names, argument meaning and error behaviour as Reachability.jl v0.7.0 (file:line under /root/reference/src):

    solve(IVP(LCS(A), X0), Options(T=0.1), op=BFFPSV18(vars=[1, 3], partition=[[1, 2], [3, 4]]))   # solve.jl:40-70
    solve(IVP(CLCCS(A, B, None, U), X0), Options(T=0.1), op=GLGM06(max_order=15))                    # GLGM06/post.jl

Julia's `:T=>0.1` pairs become keyword arguments (`δ` may also be spelled `delta` / `sampling_time`); indices in
`vars` / `partition` stay 1-based as in the reference.  Everything numerical runs in libreach_sm100a.so (device
discretisation, BFFPSV18 box recurrence, GLGM06 zonotope recurrence, reduce_order); this module only validates
options, assembles arguments and wraps the returned arrays in array-backed result containers
(`ReachSolution` -> `Flowpipe` -> `SparseReachSet` / `ReachSet`, ReachSet.jl:15-57, Flowpipe.jl:10-12,
ReachSolution.jl:12-25).  Inputs the device path does not cover raise `ReachError(3, ...)`: there is no CPU fallback.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Any, List, Optional, Sequence

import numpy as np

from ._lib import ReachError
from .backend import Context, KEEP_ZERO_GENERATORS

# ---------------------------------------------------------------------------------------------------------
# sets (the LazySets types that appear on the path)
# ---------------------------------------------------------------------------------------------------------


class LazySet:
    def dim(self) -> int:
        raise NotImplementedError

    def __mul__(self, other):          # X × Y
        return CartesianProductArray(_flatten(self) + _flatten(other))


def dim(X) -> int:
    return X.dim()


class Hyperrectangle(LazySet):
    def __init__(self, center, radius):
        self.center = np.asarray(center, dtype=np.float64).ravel()
        self.radius = np.asarray(radius, dtype=np.float64).ravel()
        assert self.center.shape == self.radius.shape
        if np.any(self.radius < 0):
            raise ValueError("radius must not be negative")

    def dim(self):
        return len(self.center)

    def low(self):
        return self.center - self.radius

    def high(self):
        return self.center + self.radius

    def __repr__(self):
        return f"{type(self).__name__}({self.center!r}, {self.radius!r})"


class BallInf(Hyperrectangle):
    def __init__(self, center, radius: float):
        c = np.asarray(center, dtype=np.float64).ravel()
        super().__init__(c, np.full(len(c), float(radius)))


class Singleton(Hyperrectangle):
    def __init__(self, element):
        c = np.asarray(element, dtype=np.float64).ravel()
        super().__init__(c, np.zeros(len(c)))


class Interval(Hyperrectangle):
    def __init__(self, lo: float, hi: float):
        assert lo <= hi
        self.lo, self.hi = float(lo), float(hi)
        super().__init__([(lo + hi) / 2], [(hi - lo) / 2])

    def low(self):
        return np.array([self.lo])

    def high(self):
        return np.array([self.hi])

    def __repr__(self):
        return f"Interval({self.lo!r}, {self.hi!r})"


class CartesianProductArray(LazySet):
    def __init__(self, array: Sequence[LazySet]):
        self.array = list(array)

    def dim(self):
        return sum(x.dim() for x in self.array)


class Zonotope(LazySet):
    def __init__(self, center, generators):
        self.center = np.asarray(center, dtype=np.float64).ravel()
        self.generators = np.asarray(generators, dtype=np.float64).reshape(len(self.center), -1)

    def dim(self):
        return len(self.center)

    def ngens(self):
        return self.generators.shape[1]


class Ball2Stub(LazySet):
    """Stand-in for a non-box LazySet (e.g. `Ball2`): accepted by the reference through LazySets' generic
    over-approximation, outside the device path here (raises instead of falling back)."""

    def __init__(self, center, radius):
        self.center = np.asarray(center, dtype=np.float64).ravel()
        self.radius = float(radius)

    def dim(self):
        return len(self.center)


class Universe(LazySet):
    def __init__(self, n):
        self.n = n

    def dim(self):
        return self.n


@dataclass
class LinearConstraint:
    a: np.ndarray
    b: float


HalfSpace = LinearConstraint


@dataclass
class SafeStatesProperty:
    """`is_contained_in(LinearConstraint(a, b))`: the states must satisfy a·x <= b (Properties/*.jl)."""
    constraint: LinearConstraint


def is_contained_in(constraint: LinearConstraint) -> SafeStatesProperty:
    return SafeStatesProperty(LinearConstraint(np.asarray(constraint.a, dtype=np.float64), float(constraint.b)))


def _flatten(X):
    return list(X.array) if isinstance(X, CartesianProductArray) else [X]


def _as_box(X, what):
    """(centre, radius) of a box-shaped set; anything else is outside the device path."""
    if isinstance(X, Hyperrectangle):
        return X.center.copy(), X.radius.copy()
    if isinstance(X, CartesianProductArray) and all(isinstance(x, Hyperrectangle) for x in X.array):
        return (np.concatenate([x.center for x in X.array]), np.concatenate([x.radius for x in X.array]))
    raise ReachError(3, f"{what} of type {type(X).__name__} is not a box (BallInf / Hyperrectangle / Interval / Singleton "
                        "or a product of them): unsupported on the sm_100a path, which has no CPU fallback")


# ---------------------------------------------------------------------------------------------------------
# systems (MathematicalSystems.jl types used by the reference; aliases Utils/systems.jl:7-21)
# ---------------------------------------------------------------------------------------------------------

@dataclass
class ConstantInput:
    U: Any


@dataclass
class VaryingInput:
    """`VaryingInput(U_1, U_2, ...)` (MathematicalSystems): a sequence of input sets.  `discretize` bloats each of them
    (discretize.jl:690-696, 729-737; `discretize_inputs` below); the reach kernels of both post operators take a
    `ConstantInput` only (reach_blocks.jl:137, GLGM06/reach.jl:32), in the reference as here."""
    Us: Sequence[Any]


@dataclass
class LinearContinuousSystem:
    A: np.ndarray


@dataclass
class ConstrainedLinearControlContinuousSystem:
    A: np.ndarray
    B: Optional[np.ndarray]
    X: Any
    U: Any


@dataclass
class InitialValueProblem:
    s: Any
    x0: Any


LCS = LinearContinuousSystem
CLCCS = ConstrainedLinearControlContinuousSystem
IVP = InitialValueProblem


def statedim(S) -> int:
    return (S.A if hasattr(S.A, "toarray") else np.asarray(S.A)).shape[0]


# ---------------------------------------------------------------------------------------------------------
# options and post operators
# ---------------------------------------------------------------------------------------------------------

_ALIASES = {"delta": "δ", "sampling_time": "δ", "time_horizon": "T", "ε": "ε"}


class Options(dict):
    """`Options(:T=>0.1, ...)` (Options/dictionary.jl): a dictionary of option name -> value."""

    def __init__(self, *args, **kw):
        super().__init__()
        for a in args:
            self.update(a)
        for k, v in kw.items():
            self[_ALIASES.get(k, k)] = v


class _Post:
    _spec: dict = {}

    def __init__(self, *args, **kw):
        specified = Options(*args, **kw)
        for k in specified:
            if k not in self._spec:
                raise ValueError(f"option '{k}' is not recognized")       # ArgumentError, Options/validation
        self.specified = specified
        self.validate()

    def __getitem__(self, k):
        return self.specified[k] if k in self.specified else self._spec[k]

    def validate(self):
        if not (self["δ"] > 0.0):
            raise ValueError(f"{self['δ']} is not a valid value for option δ")    # DomainError


def ispartition(partition) -> bool:
    """BFFPSV18.jl:12-23."""
    current = 1
    for block in partition:
        for i in block:
            if i != current:
                return False
            current += 1
    return True


class BFFPSV18(_Post):
    """`BFFPSV18 <: ContinuousPost` (BFFPSV18.jl:25-75 option table, :289-309 constructors)."""
    _spec = dict(discretization="forward", algorithm="explicit", δ=1e-2, vars=None, partition=None, sih_method="concrete",
                 exp_method="base", assume_sparse=False, lazy_inputs_interval=None, block_options=None,
                 block_options_init=None, block_options_iter=None, assume_homogeneous=False, eager_checking=True)

    def validate(self):
        super().validate()
        if self["algorithm"] not in ("explicit", "wrap"):
            raise ValueError(f"{self['algorithm']} is not a valid value for option algorithm")
        v = self["vars"]
        if v is not None and not (len(v) > 0 and all(e > 0 for e in v)):
            raise ValueError("option vars must be a non-empty vector of positive indices")
        p = self["partition"]
        if p is not None and not ispartition(p):
            raise ValueError("option partition is not a partition")
        if self["sih_method"] not in ("concrete", "lazy"):
            # symmetric_interval_hull vs the lazy SymmetricIntervalHull (discretize.jl:636-643): the same box either way
            raise ValueError(f"the method {self['sih_method']} is unknown")
        v = self["lazy_inputs_interval"]
        if not (v is None or callable(v) or (isinstance(v, int) and not isinstance(v, bool) and v >= -1)):
            # BFFPSV18.jl:51-56.  Any admissible value yields the same sets on this path: for Interval / Hyperrectangle blocks
            # the box of a lazily accumulated Minkowski sum is the sum of the boxes (axis-direction support functions add up),
            # so keeping the inputs lazy for k steps (reach.jl:114-147) cannot change the result.
            raise ValueError("option lazy_inputs_interval must be an integer >= -1 or a predicate over indices")


class GLGM06(_Post):
    """`GLGM06 <: ContinuousPost` (GLGM06/init.jl:4-29)."""
    _spec = dict(δ=1e-2, discretization="forward", sih_method="concrete", exp_method="base", max_order=10)


def compute_blocks(vars_, partition):
    """BFFPSV18.jl:413-434 (1-based)."""
    blocks, nxt, var_idx = [], 0, 0
    vars_ = list(vars_)
    for i, block in enumerate(partition, start=1):
        nxt += len(block)
        if vars_[var_idx] <= nxt:
            blocks.append(i)
            var_idx += 1
            while var_idx < len(vars_) and vars_[var_idx] <= nxt:
                var_idx += 1
            if var_idx >= len(vars_):
                break
    assert var_idx == len(vars_)
    return blocks


def compute_dimensions(partition, blocks):
    """compute_dimensions.jl:1-15 (1-based)."""
    return [d for b in blocks for d in partition[b - 1]]


# ---------------------------------------------------------------------------------------------------------
# results: array-backed containers (ReachSet.jl, Flowpipe.jl, ReachSolution.jl)
# ---------------------------------------------------------------------------------------------------------

@dataclass
class ReachSet:
    X: Any
    t_start: float
    t_end: float


@dataclass
class SparseReachSet:
    rs: ReachSet
    dimensions: List[int]


def set_of(rs):
    return rs.rs.X if isinstance(rs, SparseReachSet) else rs.X


def time_start(rs):
    return rs.rs.t_start if isinstance(rs, SparseReachSet) else rs.t_start


def time_end(rs):
    return rs.rs.t_end if isinstance(rs, SparseReachSet) else rs.t_end


class _LazyReachSets(Sequence):
    """`reachsets` vector whose elements are built from the result arrays on access (the reference allocates N
    objects up front; the arrays are what came back from the device)."""

    def __init__(self, n, make):
        self._n, self._make = n, make

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(self._n))]
        if i < 0:
            i += self._n
        if not 0 <= i < self._n:
            raise IndexError(i)
        return self._make(i)


@dataclass
class Flowpipe:
    reachsets: Any


@dataclass
class ReachSolution:
    flowpipes: List[Flowpipe]
    options: Options
    arrays: dict = field(default_factory=dict, repr=False)     # raw device results (centers/radii or handle)


@dataclass
class CheckSolution:
    satisfied: bool
    violation: int
    options: Options


def _times(N, delta):
    """t0 = t1; t1 += δ (reach_blocks.jl:200-201, GLGM06/reach.jl:20-21): sequentially accumulated."""
    t1 = np.cumsum(np.full(N, float(delta)))
    t0 = np.concatenate([[0.0], t1[:-1]])
    return t0, t1


# ---------------------------------------------------------------------------------------------------------
# solve
# ---------------------------------------------------------------------------------------------------------

_ctx_cache = {}


def _context(device=0) -> Context:
    if device not in _ctx_cache:
        _ctx_cache[device] = Context(device)
    return _ctx_cache[device]


def _normalize(problem, keep_sparse=False):
    """`normalize(system)` (Utils/normalization.jl:69-102,143-166) restricted to what the device path accepts:
    returns (A, B or None, U box (cU, rU) or None)."""
    S = problem.s
    A = S.A
    if hasattr(A, "toarray"):
        A = A if keep_sparse else A.toarray()              # Matrix(A*δ), discretize.jl:152
    else:
        A = np.asarray(A, dtype=np.float64)
    if A.ndim != 2 or A.shape[0] != A.shape[1]:
        raise ValueError("the state matrix must be square")
    if isinstance(S, LinearContinuousSystem):
        return A, None, None
    if isinstance(S, ConstrainedLinearControlContinuousSystem):
        if S.X is not None and not isinstance(S.X, Universe):
            raise ReachError(3, "state constraints (invariants) other than Universe are unsupported on the sm_100a path")
        if isinstance(S.U, VaryingInput):
            raise ReachError(3, "time-varying inputs: the reach kernels take a ConstantInput (the reference has no reach_blocks! / "
                                "reach_inhomog! method for VaryingInput either); `discretize_inputs` bloats the sequence")
        U = S.U.U if isinstance(S.U, ConstantInput) else S.U
        if U is None:
            return A, None, None
        cU, rU = _as_box(U, "input set")
        B = None if S.B is None else np.asarray(S.B, dtype=np.float64).reshape(A.shape[0], -1)
        if B is not None and B.shape[1] != len(cU):
            raise AssertionError("the input dimension does not match the number of columns of B")
        return A, B, (cU, rU)
    raise ReachError(3, f"system type {type(S).__name__} is unsupported on the sm_100a path")


def solve(problem: InitialValueProblem, options: Optional[Options] = None, op=None, device: int = 0, **kw):
    """`solve(system, options; op=BFFPSV18())` (solve.jl:40-70)."""
    opts = Options(options or {}, **kw)
    if op is None:
        op = BFFPSV18()
    if "T" not in opts:
        raise ValueError("option T (time horizon) must be specified")      # default_options.jl:72-76
    if not (opts["T"] > 0):
        raise ValueError(f"{opts['T']} is not a valid value for option T")
    opts.setdefault("mode", "reach")
    opts.setdefault("property", None)
    opts.setdefault("projection_matrix", None)
    opts.setdefault("project_reachset", False)
    if opts["mode"] == "check" and opts["property"] is None:
        raise ValueError("No property has been defined")                   # default_options.jl:91-93
    n = statedim(problem.s)
    assert dim(problem.x0) == n, "the system's state dimension and the initial set dimension must match"   # solve.jl:64
    opts["n"] = n
    if isinstance(op, BFFPSV18):
        sol = _post_bffpsv18(op, problem, opts, _context(device))
    elif isinstance(op, GLGM06):
        if op["discretization"] in ("firstorder", "nobloating"):
            # discretize(...; set_operations="zonotope") throws ArgumentError for these models (discretize.jl:105-115); running
            # them anyway would drop the bloating terms silently and return an unsound flowpipe
            raise ValueError(f"the algorithm {op['discretization']} with set operations=zonotope is not implemented")
        sol = _post_glgm06(op, problem, opts, _context(device))
    else:
        raise ReachError(3, f"post operator {type(op).__name__} is not on the sm_100a path")
    if opts["project_reachset"] and isinstance(sol, ReachSolution):
        return project(sol)                     # `:project_reachset` (BFFPSV18.jl:376-381, GLGM06/post.jl:46-49): over `:plot_vars`
    return sol


def _property_on_rows(prop, n, dims, rows):
    """`inout_map_property` (partitions.jl:33-46) for `is_contained_in(LinearConstraint(a, b))`: the normal vector restricted
    to the coordinates of the computed blocks."""
    ell = np.asarray(prop.constraint.a, dtype=np.float64)
    if len(ell) != n:
        raise AssertionError("the property has the wrong dimension")
    outside = np.setdiff1d(np.nonzero(ell)[0] + 1, dims)
    if len(outside):
        raise ValueError(f"the property involves variables {outside.tolist()} outside the computed blocks")
    return ell[rows], float(prop.constraint.b)


def _output_matrix_on_rows(M, n, rows):
    """`:projection_matrix` as output function (reach.jl:73-82): its columns on the coordinates of the computed blocks."""
    M = np.asarray(M, dtype=np.float64)
    if M.ndim != 2 or M.shape[1] != n:
        raise AssertionError("the projection matrix has the wrong number of columns")
    if np.any(np.delete(M, rows, axis=1) != 0):
        raise ValueError("the projection matrix involves variables outside the computed blocks")
    return np.ascontiguousarray(M[:, rows])


def _block_set_types(op, partition):
    """`:block_options_init` / `:block_options_iter` (BFFPSV18.jl:104-129, 436-443): only the two box types run on
    the device."""
    out = None
    for key in ("block_options_init", "block_options_iter"):
        bo = op[key] if op[key] is not None else op["block_options"]
        if bo is None:
            types = [Interval if len(b) == 1 else Hyperrectangle for b in partition]
        elif isinstance(bo, type):
            types = [bo] * len(partition)
        elif isinstance(bo, (list, tuple)):
            types = list(bo)
        elif isinstance(bo, dict):
            types = [bo.get(i + 1, Interval if len(b) == 1 else Hyperrectangle) for i, b in enumerate(partition)]
        else:
            raise ValueError(f"the `{key}` option does not accept the given input")      # DomainError
        for t, b in zip(types, partition):
            if t not in (Interval, Hyperrectangle):
                raise ReachError(3, f"block option {getattr(t, '__name__', t)} is unsupported on the sm_100a path "
                                    "(only Interval / Hyperrectangle blocks; no CPU fallback)")
            if t is Interval and len(b) != 1:
                raise ValueError("Interval blocks must be one-dimensional")
        out = types
    return out


def _post_bffpsv18(op, problem, opts, ctx):
    n = opts["n"]
    A, B, U = _normalize(problem, keep_sparse=op["exp_method"] == "lazy")
    if op["assume_homogeneous"]:
        B, U = None, None
    delta, T = float(op["δ"]), float(opts["T"])
    vars_ = list(op["vars"]) if op["vars"] is not None else list(range(1, n + 1))
    partition = [list(b) for b in op["partition"]] if op["partition"] is not None else [[i] for i in range(1, n + 1)]
    if sum(len(b) for b in partition) != n:
        raise AssertionError("the partition does not cover the state dimension")
    blocks = compute_blocks(vars_, partition)
    dims = compute_dimensions(partition, blocks)
    types = _block_set_types(op, partition)
    if op["discretization"] not in ("forward", "backward", "firstorder", "nobloating"):
        raise ValueError(f"the algorithm {op['discretization']} is unknown")          # discretize.jl:121
    if op["exp_method"] not in ("base", "sm100a", "lazy", "pade"):
        raise ValueError(f"the exponentiation method {op['exp_method']} is unknown")    # ArgumentError, discretize.jl:158
    c, r = _as_box(problem.x0, "initial set")
    N = int(math.ceil(T / delta))                                          # reach.jl:54
    rows = np.asarray(dims, dtype=np.int32) - 1
    if op["exp_method"] == "lazy":
        # ϕ = SparseMatrixExp(A*δ) (discretize.jl:153-154) -> reach_blocks!(ϕ::SparseMatrixExp, ...) (reach_blocks.jl:227-397):
        # the matrix exponential is only ever applied to blocks of vectors on the device
        Sp = ctx.sparse(A)
        try:
            D = Sp.discretize_box(delta, c, r, B if U is not None else None, None if U is None else U[0],
                                  None if U is None else U[1], algorithm=op["discretization"])
            sizes = np.asarray([len(partition[b - 1]) for b in blocks], dtype=np.int32)
            if opts["mode"] == "check":
                ell, bound = _property_on_rows(opts["property"], n, dims, rows)
                v = Sp.bffpsv18_check(delta, D["c0"], D["r0"], N, ell, bound, D["MU"], D["cU"], D["rU"], D["rpsi"], rows,
                                      eager=op["eager_checking"], block_sizes=sizes)
                return CheckSolution(v == 0, -1 if v == 0 else v, opts)
            M = opts["projection_matrix"]
            if M is not None and not opts["project_reachset"]:
                full = _output_matrix_on_rows(M, n, rows)
                C, R = Sp.bffpsv18_output_map(delta, D["c0"], D["r0"], N, full, D["MU"], D["cU"], D["rU"], D["rpsi"], rows,
                                              block_sizes=sizes)
                t0, t1 = _times(N, delta)

                def make(i):
                    return ReachSet(Hyperrectangle(C[:, i], R[:, i]), float(t0[i]), float(t1[i]))     # reach.jl:82
                return ReachSolution([Flowpipe(_LazyReachSets(N, make))], opts, dict(centers=C, radii=R))
            C, R = Sp.bffpsv18_boxes(delta, D["c0"], D["r0"], N, D["MU"], D["cU"], D["rU"], D["rpsi"], rows)
        finally:
            Sp.close()
        return _box_solution(C, R, N, delta, partition, blocks, dims, types, opts, D)
    D = ctx.discretize_box(A, delta, c, r, B if U is not None else None, None if U is None else U[0],
                           None if U is None else U[1], algorithm=op["discretization"])
    if opts["mode"] == "check":
        ell, bound = _property_on_rows(opts["property"], n, dims, rows)
        sizes = np.asarray([len(partition[b - 1]) for b in blocks], dtype=np.int32)
        v = ctx.bffpsv18_check(D["Phi"], D["c0"], D["r0"], N, ell, bound, D["MU"], D["cU"], D["rU"],
                               D["rpsi"], rows, eager=op["eager_checking"], block_sizes=sizes)
        return CheckSolution(v == 0, -1 if v == 0 else v, opts)           # BFFPSV18.jl:396-403
    t0, t1 = _times(N, delta)
    M = opts["projection_matrix"]
    if M is not None and not opts["project_reachset"]:
        sizes = np.asarray([len(partition[b - 1]) for b in blocks], dtype=np.int32)
        full = _output_matrix_on_rows(M, n, rows)
        C, R = ctx.bffpsv18_output_map(D["Phi"], D["c0"], D["r0"], N, full, D["MU"], D["cU"], D["rU"], D["rpsi"], rows,
                                       block_sizes=sizes)

        def make(i):
            return ReachSet(Hyperrectangle(C[:, i], R[:, i]), float(t0[i]), float(t1[i]))     # reach.jl:82
        return ReachSolution([Flowpipe(_LazyReachSets(N, make))], opts, dict(centers=C, radii=R))
    if op["exp_method"] == "pade":
        # `padm(A*δ)` (discretize.jl:155-156, Expokit's Padé approximant for a sparse A): the same matrix e^{Aδ} to rounding --
        # the device computes it with its own scaling-and-squaring scheme -- used in SPARSE form, i.e. through the sparse twin
        # of reach_blocks! (reach_blocks.jl:47-132): decoupled subsystems skip their zero blocks on the device
        import scipy.sparse as sp
        C, R = ctx.bffpsv18_boxes_csc(sp.csc_matrix(D["Phi"]), D["c0"], D["r0"], N, D["MU"], D["cU"], D["rU"], D["rpsi"], rows)
    else:
        C, R = ctx.bffpsv18_boxes(D["Phi"], D["c0"], D["r0"], N, D["MU"], D["cU"], D["rU"], D["rpsi"], rows)
    return _box_solution(C, R, N, delta, partition, blocks, dims, types, opts, D)


def _box_solution(C, R, N, delta, partition, blocks, dims, types, opts, D):
    """The `SparseReachSet{CartesianProductArray}` flowpipe of reach.jl:79 over the (centers, radii) arrays."""
    t0, t1 = _times(N, delta)
    offs = np.concatenate([[0], np.cumsum([len(partition[b - 1]) for b in blocks])])

    def make(i):
        arr = []
        for j, b in enumerate(blocks):
            lo, hi = offs[j], offs[j + 1]
            if types[b - 1] is Interval:
                arr.append(Interval(C[lo, i] - R[lo, i], C[lo, i] + R[lo, i]))
            else:
                arr.append(Hyperrectangle(C[lo:hi, i], R[lo:hi, i]))
        return SparseReachSet(ReachSet(CartesianProductArray(arr), float(t0[i]), float(t1[i])), list(dims))   # reach.jl:79
    return ReachSolution([Flowpipe(_LazyReachSets(N, make))], opts, dict(centers=C, radii=R, dimensions=dims, discretized=D))


def discretize_zonotope(ctx: Context, A, delta, c, r, B=None, U=None, algorithm="forward", remove_zero_generators=True):
    """`discretize(...; set_operations="zonotope")` (discretize.jl:705-743) for a box X0: Φ, Φ₂(|A|), Einit, Eψ0 AND the
    concrete zonotope arithmetic (convert, linear_map, minkowski_sum, overapproximate(CH(Z1, Z2), Zonotope), the
    constructor's zero-generator removal) run on the device behind `reach_discretize_zonotope`.
    Returns (Phi, (c0, G0), (cV, GV) or None).  `firstorder` / `nobloating` raise like the reference's ArgumentError
    (discretize.jl:105-115)."""
    if algorithm in ("firstorder", "nobloating"):
        raise ValueError(f"the algorithm {algorithm} with set operations=zonotope is not implemented")
    if algorithm not in ("forward", "backward"):
        raise ValueError(f"the algorithm {algorithm} is unknown")                      # discretize.jl:121
    D = ctx.discretize_zonotope(A, delta, c, r, B if U is not None else None, None if U is None else U[0],
                                None if U is None else U[1], algorithm=algorithm, remove_zero_generators=remove_zero_generators)
    V = None if U is None else (D["cV"], D["GV"])
    return D["Phi"], (D["c0"], D["G0"]), V


def discretize_inputs(ctx: Context, A, delta, B, U: "VaryingInput"):
    """The `VaryingInput` branch of `discretize` (discretize.jl:690-696): returns (MU = δB, [(cU_k, rU_k, rpsi_k)]), i.e.
    Ud[k] = MU * Box(cU_k, rU_k) ⊕ Box(0, rpsi_k) for every input set of the sequence, all k in one device call."""
    boxes = [_as_box(Uk, "input set") for Uk in U.Us]
    cUs = np.stack([b[0] for b in boxes], axis=1)
    rUs = np.stack([b[1] for b in boxes], axis=1)
    D = ctx.discretize_inputs_box(A, delta, B, cUs, rUs)
    return D["MU"], [(cUs[:, k], rUs[:, k], D["rpsi"][:, k]) for k in range(len(boxes))]


def _post_glgm06(op, problem, opts, ctx):
    n = opts["n"]
    A, B, U = _normalize(problem)
    delta, T = float(op["δ"]), float(opts["T"])
    max_order = int(op["max_order"])
    N = int(round(T / delta))                                                # post.jl:12
    if N < 1:
        raise ValueError("the time horizon is shorter than half a time step")
    c, r = _as_box(problem.x0, "initial set")
    if op["discretization"] in ("firstorder", "nobloating"):                 # ArgumentError, discretize.jl:105-115
        raise ValueError(f"the algorithm {op['discretization']} with set operations=zonotope is not implemented")
    Phi, (c0, G0), V = discretize_zonotope(ctx, A, delta, c, r, B, U, op["discretization"])
    G0red, _ = ctx.reduce_order(c0, G0, max_order)                           # post.jl:20
    fp = ctx.glgm06(Phi, c0, G0red, N, max_order, None if V is None else V[0], None if V is None else V[1], 0)
    fp.run()
    p = fp.ngens()
    t0, t1 = _times(N, delta)

    def make(i):
        ck, Gk = fp.step(i + 1, int(p[i]))
        return ReachSet(Zonotope(ck, Gk), float(t0[i]), float(t1[i]))       # post.jl:27
    return ReachSolution([Flowpipe(_LazyReachSets(N, make))], opts, dict(flowpipe=fp, ngens=p, Phi=Phi, Omega0=(c0, G0red), V=V, times=(t0, t1)))


def _projection_rows(vars_, n, M, T):
    """The projection matrix `project_reach` assembles (project_reach.jl:53-86): rows over the state variables; with
    `vars_[0] == 0` (time) only the y row.  M: optional 1 x n `:projection_matrix`, T: optional `:transformation_matrix`."""
    xaxis, yaxis = vars_
    got_time = xaxis == 0
    if xaxis < 0:
        raise ValueError(f"value {xaxis} for x variable not allowed")          # DomainError, project_reach.jl:27-28
    if yaxis <= 0:
        raise ValueError(f"value {yaxis} for y variable not allowed")          # :29-30
    if M is None:
        P = np.zeros((1 if got_time else 2, n))
        if got_time:
            P[0, yaxis - 1] = 1.0
        else:
            P[0, xaxis - 1] = 1.0
            P[1, yaxis - 1] = 1.0
    else:
        M = np.asarray(M, dtype=np.float64).reshape(-1, n)
        assert M.shape[0] == 1, "currently we only support one-dimensional projection matrices"      # :64-65
        if got_time:
            P = M[:1].copy()
        else:
            P = np.zeros((2, n))
            P[1] = M[0]
            P[0, xaxis - 1] = 1.0
    if T is not None:
        P = P @ np.asarray(T, dtype=np.float64)                                   # :78-86 (state part of [S 0; 0 1])
    return P, got_time


def project(sol: ReachSolution, vars_: Optional[Sequence[int]] = None):
    """`project(sol)` (project_reach.jl:18-99) with the default `set_type_proj = Hyperrectangle`: 2-D boxes over
    `plot_vars` (0 = time).  Zonotope flowpipes (GLGM06) are projected ON THE DEVICE (`reach_flowpipe_project`), only the
    2 x N boxes come back; decomposed box flowpipes (BFFPSV18) are already per-coordinate boxes on the host."""
    vars_ = list(vars_ if vars_ is not None else sol.options.get("plot_vars", [0, 1]))
    assert len(vars_) == 2, "we only support projection to two dimensions"
    rs = sol.flowpipes[0].reachsets
    fp = sol.arrays.get("flowpipe")
    if fp is not None:
        n = fp.n
        P, got_time = _projection_rows(vars_, n, sol.options.get("projection_matrix"), sol.options.get("transformation_matrix"))
        Cp, Rp = fp.project(P)
        N = Cp.shape[1]
        t0, t1 = sol.arrays["times"]
        out = []
        for i in range(N):
            if got_time:
                c = np.array([(t0[i] + t1[i]) / 2, Cp[0, i]])
                r = np.array([(t1[i] - t0[i]) / 2, Rp[0, i]])
            else:
                c, r = Cp[:, i].copy(), Rp[:, i].copy()
            out.append(ReachSet(Hyperrectangle(c, r), float(t0[i]), float(t1[i])))
        return ReachSolution([Flowpipe(out)], sol.options, dict(centers=Cp, radii=Rp, projection=P))
    C, R, dims = sol.arrays.get("centers"), sol.arrays.get("radii"), sol.arrays.get("dimensions")
    if dims is None:
        raise ReachError(3, "project: only decomposed box flowpipes and zonotope flowpipes are supported on this path")
    out = []
    for i in range(len(rs)):
        lo, hi = [], []
        for v in vars_:
            if v == 0:
                lo.append(time_start(rs[i])); hi.append(time_end(rs[i]))
            else:
                j = dims.index(v)
                lo.append(C[j, i] - R[j, i]); hi.append(C[j, i] + R[j, i])
        lo, hi = np.array(lo), np.array(hi)
        out.append(ReachSet(Hyperrectangle((lo + hi) / 2, (hi - lo) / 2), time_start(rs[i]), time_end(rs[i])))
    return ReachSolution([Flowpipe(out)], sol.options)
