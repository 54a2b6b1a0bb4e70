"""GPU: the host-side mirror of the reference interface (`solve`, `BFFPSV18`, `GLGM06`, options, result containers),
written after /root/reference/test/Reachability/solve_continuous.jl.  Values are checked against the oracle."""
import numpy as np
import pytest

from oracle import reach_oracle as orc
from reachability_jl_b200 import ReachError
from reachability_jl_b200.api import (BFFPSV18, CLCCS, GLGM06, IVP, LCS, BallInf, Ball2Stub, ConstantInput, Hyperrectangle,
                                      Interval, LinearConstraint, Options, Singleton, dim, is_contained_in, project, set_of,
                                      solve, time_end, time_start)
from tests.util import assert_close

pytestmark = pytest.mark.gpu

A = np.array([[0.0509836, 0.168159, 0.95246, 0.33644],
              [0.42377, 0.67972, 0.129232, 0.126662],
              [0.518654, 0.981313, 0.489854, 0.588326],
              [0.38318, 0.616014, 0.518412, 0.778765]])
B = np.array([[0.866688, 0.771231], [0.065935, 0.349839], [0.109643, 0.904222], [0.292474, 0.227857]])
X0 = BallInf(np.ones(4), 0.1)


def test_default_options_all_variables():                                   # solve_continuous.jl:15-17
    s = solve(IVP(LCS(A), X0), T=0.1)
    rs = s.flowpipes[0].reachsets
    assert len(rs) == 10 and dim(set_of(rs[0])) == 4 and isinstance(s.options, Options)
    assert all(isinstance(b, Interval) for b in set_of(rs[0]).array)        # default 1-D partition -> Interval blocks
    D = orc.discretize_box(A, orc.Box(np.ones(4), np.full(4, 0.1)), 0.01)
    Co, Ro = orc.bffpsv18_reach(D, 10)
    for k in (0, 4, 9):
        lo = np.array([b.lo for b in set_of(rs[k]).array])
        hi = np.array([b.hi for b in set_of(rs[k]).array])
        assert_close(lo, Co[k] - Ro[k], "lo")
        assert_close(hi, Co[k] + Ro[k], "hi")
    assert time_start(rs[0]) == 0.0 and abs(time_end(rs[9]) - 0.1) < 1e-12 and time_start(rs[3]) == time_end(rs[2])


def test_vars_and_custom_partition():                                       # :20-29
    s = solve(IVP(LCS(A), X0), Options(T=0.1), op=BFFPSV18(vars=[1, 3], partition=[[1, 2], [3, 4]]))
    rs0 = s.flowpipes[0].reachsets[0]
    assert dim(set_of(rs0)) == 4 and rs0.dimensions == [1, 2, 3, 4]
    assert all(type(b) is Hyperrectangle for b in set_of(rs0).array)
    s = solve(IVP(LCS(A), X0), Options(T=0.1), op=BFFPSV18(vars=[1, 3]))
    rs0 = s.flowpipes[0].reachsets[0]
    assert dim(set_of(rs0)) == 2 and rs0.dimensions == [1, 3]


def test_dimension_mismatch_is_an_assertion_error():                        # :32
    with pytest.raises(AssertionError):
        solve(IVP(LCS(A), BallInf(np.ones(3), 0.1)), T=0.1)


def test_option_validation():
    with pytest.raises(ValueError):
        BFFPSV18(δ=-1.0)
    with pytest.raises(ValueError):
        BFFPSV18(partition=[[1, 3], [2, 4]])
    with pytest.raises(ValueError):
        BFFPSV18(bogus=1)
    with pytest.raises(ValueError):
        solve(IVP(LCS(A), X0), Options())                                   # T is mandatory
    with pytest.raises(ValueError):
        solve(IVP(LCS(A), X0), Options(T=0.1, mode="check"))                # check mode needs a property
    with pytest.raises(ValueError):
        solve(IVP(LCS(A), X0), Options(T=0.1), op=BFFPSV18(discretization="sideways"))


def test_check_mode_known_answers():                                        # :49-75
    prop = is_contained_in(LinearConstraint([1., 1., 0., 0.], 2.))
    s = solve(IVP(LCS(A), X0), Options(T=0.1, mode="check", property=prop), op=BFFPSV18(vars=[1, 2], partition=[[1, 2], [3, 4]]))
    assert s.violation == 1 and not s.satisfied
    s = solve(IVP(LCS(A), X0), Options(T=0.1, mode="check", property=prop), op=BFFPSV18(vars=[1, 2, 3], partition=[[1, 2], [3, 4]]))
    assert s.violation == 1
    prop = is_contained_in(LinearConstraint([1., -1., 0., 0.], 2.))
    s = solve(IVP(LCS(A), X0), Options(T=0.1, mode="check", property=prop), op=BFFPSV18(vars=[1, 2, 3], partition=[[1, 2], [3, 4]]))
    assert s.violation == -1 and s.satisfied


def test_projection_matrix_output_function():                               # :82-83
    M = np.array([[1.0, 0.0, 0.0, 0.0], [0.0, 1.0, 0.0, 0.0]])
    s = solve(IVP(LCS(A), X0), Options(T=1.0, projection_matrix=M), op=BFFPSV18(δ=0.1))
    rs = s.flowpipes[0].reachsets
    assert len(rs) == 10 and dim(set_of(rs[0])) == 2
    D = orc.discretize_box(A, orc.Box(np.ones(4), np.full(4, 0.1)), 0.1)
    Co, Ro = orc.bffpsv18_output_map(D, 10, M, np.arange(4), [1, 1, 1, 1])
    assert_close(np.array([set_of(r).center for r in rs]), Co, "centres")
    assert_close(np.array([set_of(r).radius for r in rs]), Ro, "radii")


def test_interval_block_option_and_project():                               # :116-120, :41-43
    s = solve(IVP(LCS(A), X0), Options(T=0.1), op=BFFPSV18(vars=[1, 3], partition=[[i] for i in range(1, 5)], block_options=Interval))
    assert all(isinstance(b, Interval) for b in set_of(s.flowpipes[0].reachsets[2]).array)
    ps = project(s, [0, 1])
    assert dim(set_of(ps.flowpipes[0].reachsets[0])) == 2


def test_affine_system_with_inputs():                                       # :157-163
    U = Interval(0.99, 1.01) * Interval(0.99, 1.01)
    for inputs in (ConstantInput(U), U):
        s = solve(IVP(CLCCS(A, B, None, inputs), X0), T=0.1)
        rs = s.flowpipes[0].reachsets
        D = orc.discretize_box(A, orc.Box(np.ones(4), np.full(4, 0.1)), 0.01, B, orc.Box(np.ones(2), np.full(2, 0.01)))
        Co, Ro = orc.bffpsv18_reach(D, 10)
        lo = np.array([b.lo for b in set_of(rs[9]).array])
        assert_close(lo, Co[9] - Ro[9], "lo")
    s2 = solve(IVP(CLCCS(np.array([[0.0, 1.0], [0.0, 0.0]]), np.array([[0.0], [-1.0]]), None, Singleton([1.0])),
                   BallInf(np.zeros(2), 0.1)), T=1.0)
    assert len(s2.flowpipes[0].reachsets) == 100


def test_glgm06_mirror():                                                   # :186-196
    for op, kw in ((GLGM06(), {}), (GLGM06(max_order=15), {})):
        s = solve(IVP(LCS(A), X0), Options(T=0.1), op=op)
        rs = s.flowpipes[0].reachsets
        assert len(rs) == 10
        Dz = orc.discretize_zonotope(A, orc.Box(np.ones(4), np.full(4, 0.1)), 0.01)
        R = orc.glgm06_reach(Dz, 10, op["max_order"])
        for k in (0, 5, 9):
            assert_close(set_of(rs[k]).center, R[k].c, "centre")
            assert_close(set_of(rs[k]).generators, R[k].G, "generators")
    U = Interval(0.99, 1.01) * Interval(0.99, 1.01)
    s = solve(IVP(CLCCS(A, B, None, U), X0), Options(T=0.1), op=GLGM06())
    Dz = orc.discretize_zonotope(A, orc.Box(np.ones(4), np.full(4, 0.1)), 0.01, B, orc.Box(np.ones(2), np.full(2, 0.01)))
    R = orc.glgm06_reach(Dz, 10, 10)
    rs = s.flowpipes[0].reachsets
    for k in (0, 1, 9):
        assert_close(set_of(rs[k]).center, R[k].c, "centre")
        assert_close(set_of(rs[k]).generators, R[k].G, "generators")


def test_glgm06_project_on_the_device():                                   # project_reach.jl:18-99 on a zonotope flowpipe
    U = Interval(0.99, 1.01) * Interval(0.99, 1.01)
    s = solve(IVP(CLCCS(A, B, None, U), X0), Options(T=0.1), op=GLGM06())
    Dz = orc.discretize_zonotope(A, orc.Box(np.ones(4), np.full(4, 0.1)), 0.01, B, orc.Box(np.ones(2), np.full(2, 0.01)))
    R = orc.glgm06_reach(Dz, 10, 10)
    p13 = project(s, [1, 3]).flowpipes[0].reachsets             # two state variables
    pt2 = project(s, [0, 2]).flowpipes[0].reachsets             # time x state variable
    for k in (0, 4, 9):
        rad = np.abs(R[k].G).sum(axis=1)
        assert_close(set_of(p13[k]).center, R[k].c[[0, 2]], "centre (1,3)")
        assert_close(set_of(p13[k]).radius, rad[[0, 2]], "radius (1,3)")
        assert_close(set_of(pt2[k]).center, np.array([0.01 * k + 0.005, R[k].c[1]]), "centre (t,2)")
        assert_close(set_of(pt2[k]).radius, np.array([0.005, rad[1]]), "radius (t,2)")
    with pytest.raises(ValueError):
        project(s, [1, 0])                                         # DomainError: y variable must be positive (:29-30)


def test_unsupported_inputs_raise_instead_of_falling_back():
    with pytest.raises(ReachError) as ei:
        solve(IVP(LCS(A), Ball2Stub(np.ones(4), 0.1)), Options(T=0.1), op=GLGM06())     # :199-200 needs LazySets on the host
    assert ei.value.code == 3


def test_lazy_exp_method_runs_the_sparse_path():                            # BFFPSV18(:exp_method => "lazy"), discretize.jl:153-154
    import scipy.sparse as sp
    g = np.random.default_rng(3)
    n = 12
    Ad = np.diag(-g.uniform(0.5, 2.0, n)) + np.diag(g.standard_normal(n - 1) * 0.3, 1) + np.diag(g.standard_normal(n - 1) * 0.3, -1)
    Bd = g.standard_normal((n, 1))
    prob = lambda Amat: IVP(CLCCS(Amat, Bd, None, Interval(0.9, 1.1)), BallInf(np.ones(n), 0.05))
    op = lambda method: BFFPSV18(δ=0.02, vars=[1, 4, 12], exp_method=method)
    dense = solve(prob(Ad), Options(T=0.4), op=op("base"))
    lazy = solve(prob(sp.csc_matrix(Ad)), Options(T=0.4), op=op("lazy"))
    lazy_dense_input = solve(prob(Ad), Options(T=0.4), op=op("lazy"))
    assert_close(lazy.arrays["centers"], dense.arrays["centers"], "centres")
    assert_close(lazy.arrays["radii"], dense.arrays["radii"], "radii")
    assert np.array_equal(lazy_dense_input.arrays["radii"], lazy.arrays["radii"])
    assert len(lazy.flowpipes[0].reachsets) == 20 and dim(set_of(lazy.flowpipes[0].reachsets[3])) == 3
    with pytest.raises(ValueError):
        solve(prob(Ad), Options(T=0.4), op=op("krylov"))                  # ArgumentError: unknown method (discretize.jl:158)


def test_reference_option_variants_of_solve_continuous():                   # solve_continuous.jl:90-121
    """The option variants the reference's own test file runs through `solve`: lazy / concrete sih, a lazy matrix exponential
    on a sparse A with and without :assume_sparse, lazy input handling, Interval blocks with a lazy exponential.  For box
    blocks they all describe the same flowpipe; every variant must run and agree with the default one."""
    import scipy.sparse as sp
    base = solve(IVP(LCS(A), X0), Options(T=0.1), op=BFFPSV18(vars=[1, 3], partition=[[1, 2], [3, 4]]))
    variants = [
        (A, dict(sih_method="lazy")),                                         # :91-93
        (A, dict(sih_method="concrete")),                                     # :95-97
        (sp.csc_matrix(A), dict(exp_method="lazy", assume_sparse=False)),      # :99-102
        (sp.csc_matrix(A), dict(exp_method="lazy", assume_sparse=True)),       # :104-107
        (A, dict(lazy_inputs_interval=lambda k: False)),                      # :109-112
    ]
    for Amat, kw in variants:
        s = solve(IVP(LCS(Amat), X0), Options(T=0.1), op=BFFPSV18(vars=[1, 3], partition=[[1, 2], [3, 4]], **kw))
        assert len(s.flowpipes[0].reachsets) == 10 and dim(set_of(s.flowpipes[0].reachsets[0])) == 4
        assert_close(s.arrays["centers"], base.arrays["centers"], f"centres {kw}")
        assert_close(s.arrays["radii"], base.arrays["radii"], f"radii {kw}")
    s = solve(IVP(LCS(sp.csc_matrix(A)), X0), Options(T=0.1),
              op=BFFPSV18(vars=[1, 3], partition=[[i] for i in range(1, 5)], exp_method="lazy", block_options=Interval))   # :116-121
    rs = s.flowpipes[0].reachsets
    assert len(rs) == 10 and dim(set_of(rs[0])) == 2
    dense = solve(IVP(LCS(A), X0), Options(T=0.1), op=BFFPSV18(vars=[1, 3], partition=[[i] for i in range(1, 5)], block_options=Interval))
    assert_close(s.arrays["centers"], dense.arrays["centers"], "Interval blocks, lazy exponential: centres")
    assert_close(s.arrays["radii"], dense.arrays["radii"], "Interval blocks, lazy exponential: radii")


def test_check_mode_and_output_function_with_a_lazy_exponential():
    """The check-mode known answers of solve_continuous.jl:49-75 and an output function, through exp_method = "lazy"."""
    import scipy.sparse as sp
    As = sp.csc_matrix(A)
    for prop_a, expect_violation in ((np.array([1.0, 1.0, 0.0, 0.0]), True), (np.array([1.0, -1.0, 0.0, 0.0]), False)):
        prop = is_contained_in(LinearConstraint(prop_a, 2.0))
        for method in ("base", "lazy"):
            s = solve(IVP(LCS(As if method == "lazy" else A), X0), Options(T=0.1, mode="check", property=prop),
                      op=BFFPSV18(vars=[1, 2, 3], partition=[[1, 2], [3, 4]], exp_method=method))
            assert s.satisfied == (not expect_violation), (method, prop_a)
            assert (s.violation == 1) if expect_violation else (s.violation == -1)
    M = np.array([[1.0, 0.5, 0.0, 0.0], [0.0, 0.0, 1.0, -1.0]])
    kw = dict(vars=[1, 2, 3, 4], partition=[[1, 2], [3, 4]])
    sd = solve(IVP(LCS(A), X0), Options(T=0.1, projection_matrix=M), op=BFFPSV18(**kw))
    sl = solve(IVP(LCS(As), X0), Options(T=0.1, projection_matrix=M), op=BFFPSV18(exp_method="lazy", **kw))
    assert_close(sl.arrays["centers"], sd.arrays["centers"], "output function centres")
    assert_close(sl.arrays["radii"], sd.arrays["radii"], "output function radii")


def test_project_reachset_option_projects_inside_solve():                   # BFFPSV18.jl:376-381, GLGM06/post.jl:46-49
    for op in (GLGM06(), BFFPSV18(vars=[1, 3], partition=[[1, 2], [3, 4]])):
        s = solve(IVP(LCS(A), X0), Options(T=0.1, project_reachset=True, plot_vars=[1, 3]), op=op)
        plain = solve(IVP(LCS(A), X0), Options(T=0.1, plot_vars=[1, 3]), op=op)
        want = project(plain).flowpipes[0].reachsets
        got = s.flowpipes[0].reachsets
        assert len(got) == 10 and dim(set_of(got[0])) == 2
        for k in (0, 9):
            assert_close(set_of(got[k]).center, set_of(want[k]).center, "centre")
            assert_close(set_of(got[k]).radius, set_of(want[k]).radius, "radius")


def test_varying_input_is_discretised_but_not_propagated():               # discretize.jl:690-696; reach_blocks.jl:137
    from reachability_jl_b200 import api
    ctx = api._context(0)
    A = np.array([[-1.0, 0.5], [0.0, -2.0]])
    B = np.array([[1.0], [0.5]])
    U = api.VaryingInput([Interval(0.9, 1.1), Interval(-0.2, 0.4), Singleton([0.3])])
    MU, seq = api.discretize_inputs(ctx, A, 0.05, B, U)
    assert np.allclose(MU, 0.05 * B) and len(seq) == 3
    assert np.all(seq[0][2] > 0) and np.all(seq[2][2] > 0)                 # E_psi_k bloats every step's input
    with pytest.raises(ReachError) as ei:
        solve(IVP(CLCCS(A, B, None, U), BallInf(np.zeros(2), 0.1)), Options(T=0.1))
    assert ei.value.code == 3


def test_pade_exp_method_runs_the_sparse_twin():                            # BFFPSV18(:exp_method => "pade"), discretize.jl:155-156
    """`padm(A*δ)` keeps ϕ sparse, so the reference takes reach_blocks!(ϕ::SparseMatrixCSC, ...) (reach_blocks.jl:47-132); the
    mirror routes the flowpipe through reach_bffpsv18_boxes_csc: decoupled subsystems skip their zero blocks, same sets."""
    import scipy.sparse as sp
    g = np.random.default_rng(8)
    n = 14
    Ad = np.zeros((n, n))
    for i in range(0, n, 2):                       # seven decoupled damped oscillators, states interleaved by a permutation
        w = g.uniform(1.0, 6.0)
        Ad[i:i + 2, i:i + 2] = [[-0.1, w], [-w, -0.1]]
    perm = g.permutation(n)
    Ad = Ad[np.ix_(perm, perm)]
    Bd = g.standard_normal((n, 1))
    prob = lambda Amat: IVP(CLCCS(Amat, Bd, None, Interval(0.9, 1.1)), BallInf(np.ones(n), 0.05))
    op = lambda method: BFFPSV18(δ=0.02, vars=[1, 4, 13], exp_method=method)
    from reachability_jl_b200 import api
    dense = solve(prob(Ad), Options(T=0.6), op=op("base"))
    pade = solve(prob(sp.csc_matrix(Ad)), Options(T=0.6), op=op("pade"))
    info = api._context(0).csc_info()
    assert info["sparse_path"] and info["max_component"] == 2
    assert_close(pade.arrays["centers"], dense.arrays["centers"], "centres")
    assert_close(pade.arrays["radii"], dense.arrays["radii"], "radii")
    assert len(pade.flowpipes[0].reachsets) == 30 and dim(set_of(pade.flowpipes[0].reachsets[7])) == 3
