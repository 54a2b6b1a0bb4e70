"""CPU: the host-side mirror's option handling and bookkeeping (no device needed) -- option tables and validation
(BFFPSV18.jl:25-129, GLGM06/init.jl:4-29), `compute_blocks` / `compute_dimensions` (BFFPSV18.jl:413-434,
compute_dimensions.jl:1-15), set helpers, time stamps, and the absence of any CPU fallback."""
import numpy as np
import pytest

from oracle import reach_oracle as orc
from reachability_jl_b200 import ReachError
from reachability_jl_b200 import api
from reachability_jl_b200.api import (BFFPSV18, GLGM06, IVP, LCS, BallInf, Ball2Stub, CartesianProductArray, Hyperrectangle,
                                      Interval, Options, Singleton, dim, ispartition)


def test_option_tables_and_aliases():
    op = BFFPSV18(delta=0.05, vars=[1, 3], partition=[[1, 2], [3, 4]])
    assert op["δ"] == 0.05 and op["discretization"] == "forward" and op["algorithm"] == "explicit"
    assert BFFPSV18(sampling_time=0.2)["δ"] == 0.2
    assert GLGM06()["max_order"] == 10 and GLGM06()["δ"] == 1e-2
    assert Options(time_horizon=3.0)["T"] == 3.0
    for bad in (dict(δ=0.0), dict(algorithm="implicit"), dict(vars=[]), dict(vars=[0, 1]), dict(partition=[[2], [1]]), dict(nope=1)):
        with pytest.raises(ValueError):
            BFFPSV18(**bad)
    with pytest.raises(ValueError):
        GLGM06(order=3)


def test_ispartition_and_blocks_match_the_oracle():
    assert ispartition([[1, 2], [3], [4, 5, 6]]) and not ispartition([[1, 3], [2]]) and not ispartition([[2, 3]])
    part = [[1, 2], [3, 4], [5], [6, 7, 8]]
    for vars_ in ([1], [3], [1, 3], [5, 8], [2, 4, 6], list(range(1, 9))):
        assert api.compute_blocks(vars_, part) == orc.compute_blocks(vars_, part)
        b = api.compute_blocks(vars_, part)
        assert api.compute_dimensions(part, b) == orc.compute_dimensions(part, b)


def test_sets_and_products():
    X = Interval(0.99, 1.01) * Interval(-1.0, 1.0) * BallInf([0.0, 2.0], 0.5)
    assert isinstance(X, CartesianProductArray) and dim(X) == 4
    c, r = api._as_box(X, "set")
    assert np.allclose(c, [1.0, 0.0, 0.0, 2.0]) and np.allclose(r, [0.01, 1.0, 0.5, 0.5])
    assert np.array_equal(Singleton([1.0, 2.0]).radius, [0.0, 0.0])
    with pytest.raises(ValueError):
        Hyperrectangle([0.0], [-1.0])
    with pytest.raises(ReachError) as ei:
        api._as_box(Ball2Stub([0.0, 0.0], 1.0), "initial set")
    assert ei.value.code == 3 and "no CPU fallback" in str(ei.value)


def test_time_stamps_accumulate_like_the_reference():
    t0, t1 = api._times(5, 0.1)                       # t0 = t1; t1 += δ  (reach_blocks.jl:200-201)
    acc, ref0, ref1 = 0.1, [0.0], [0.1]
    for _ in range(4):
        ref0.append(acc)
        acc += 0.1
        ref1.append(acc)
    assert list(t0) == ref0 and list(t1) == ref1


def test_solve_argument_errors_come_before_any_device_work():
    A = np.eye(3)
    with pytest.raises(ValueError):
        api.solve(IVP(LCS(A), BallInf(np.ones(3), 0.1)), Options())                 # no time horizon
    with pytest.raises(ValueError):
        api.solve(IVP(LCS(A), BallInf(np.ones(3), 0.1)), Options(T=-1.0))
    with pytest.raises(ValueError):
        api.solve(IVP(LCS(A), BallInf(np.ones(3), 0.1)), Options(T=1.0, mode="check"))
    with pytest.raises(AssertionError):
        api.solve(IVP(LCS(A), BallInf(np.ones(2), 0.1)), Options(T=1.0))            # solve.jl:64


def test_block_option_rules():
    part = [[1], [2, 3]]
    assert api._block_set_types(BFFPSV18(), part) == [Interval, Hyperrectangle]       # BFFPSV18.jl:436-443
    assert api._block_set_types(BFFPSV18(block_options=Hyperrectangle), part) == [Hyperrectangle, Hyperrectangle]
    with pytest.raises(ValueError):
        api._block_set_types(BFFPSV18(block_options=Interval), part)                  # Interval needs a 1-D block
    with pytest.raises(ReachError):
        api._block_set_types(BFFPSV18(block_options=Ball2Stub), part)                 # not on the device path
    with pytest.raises(ValueError):
        api._block_set_types(BFFPSV18(block_options=-1.0), part)                      # DomainError in the reference


def test_projection_matrix_assembly_follows_project_reach():
    """`project_reach` (project_reach.jl:53-86): rows over the state variables; time (0) drops out of the matrix."""
    P, got_time = api._projection_rows([1, 3], 4, None, None)
    assert not got_time and np.array_equal(P, np.array([[1.0, 0, 0, 0], [0, 0, 1.0, 0]]))
    P, got_time = api._projection_rows([0, 2], 4, None, None)
    assert got_time and np.array_equal(P, np.array([[0, 1.0, 0, 0]]))
    M = np.array([[1.0, -1.0, 0.5, 0.0]])
    P, got_time = api._projection_rows([0, 2], 4, M, None)          # 1 x n projection matrix against time (:69-71)
    assert got_time and np.array_equal(P, M)
    P, _ = api._projection_rows([2, 3], 4, M, None)                 # against a state variable (:72-76)
    assert np.array_equal(P, np.array([[0, 1.0, 0, 0], M[0]]))
    T = np.diag([2.0, 3.0, 4.0, 5.0])
    P, _ = api._projection_rows([1, 2], 4, None, T)                 # :78-86
    assert np.array_equal(P, np.array([[2.0, 0, 0, 0], [0, 3.0, 0, 0]]))
    with pytest.raises(ValueError):
        api._projection_rows([-1, 2], 4, None, None)                # DomainError (:27-28)
    with pytest.raises(ValueError):
        api._projection_rows([1, 0], 4, None, None)                 # DomainError (:29-30)
    with pytest.raises(AssertionError):
        api._projection_rows([1, 2], 4, np.ones((2, 4)), None)      # only one-dimensional projection matrices (:64-65)


def test_csc_fields_follow_julia_sparse_matrix_csc():
    """What crosses the ABI for a sparse A is exactly `A.colptr`, `A.rowval`, `A.nzval` of Julia's SparseMatrixCSC."""
    import scipy.sparse as sp
    from reachability_jl_b200.backend import _csc_fields
    A = np.array([[1.0, 0, 2.0], [0, 0, 3.0], [4.0, 5.0, 0]])
    for form in (A, sp.csc_matrix(A), sp.csr_matrix(A), sp.coo_matrix(A)):
        n, colptr, rowval, nzval = _csc_fields(form)
        assert n == 3 and colptr.dtype == np.int64 and rowval.dtype == np.int64
        assert colptr.tolist() == [1, 3, 4, 6] and rowval.tolist() == [1, 3, 3, 1, 2] and nzval.tolist() == [1.0, 4.0, 5.0, 2.0, 3.0]
    n, colptr, rowval, nzval = _csc_fields(np.zeros((2, 2)))
    assert colptr.tolist() == [1, 1, 1] and len(rowval) == 0 and len(nzval) == 0


def test_discretization_and_input_options_are_validated():      # BFFPSV18.jl:43-56, discretize.jl:636-643
    for ok in (dict(sih_method="lazy"), dict(sih_method="concrete"), dict(lazy_inputs_interval=-1), dict(lazy_inputs_interval=0),
               dict(lazy_inputs_interval=5), dict(lazy_inputs_interval=lambda k: False), dict(exp_method="lazy", assume_sparse=True)):
        BFFPSV18(**ok)
    for bad in (dict(sih_method="eager"), dict(lazy_inputs_interval=-2), dict(lazy_inputs_interval="never")):
        with pytest.raises(ValueError):
            BFFPSV18(**bad)


def test_glgm06_rejects_discretisations_without_zonotope_operations():
    """discretize.jl:105-115: `firstorder` / `nobloating` with set_operations="zonotope" throw ArgumentError in the reference;
    the mirror raises before any device work instead of silently dropping the bloating terms (unsound flowpipe)."""
    prob = IVP(LCS(-np.eye(3)), BallInf(np.ones(3), 0.1))
    for alg in ("firstorder", "nobloating"):
        with pytest.raises(ValueError, match="with set operations=zonotope is not implemented"):
            api.solve(prob, Options(T=0.1), op=GLGM06(discretization=alg))
