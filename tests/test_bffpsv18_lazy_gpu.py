"""GPU parity: the lazy-state variants of the BFFPSV18 sweep -- check mode (check_blocks.jl:105-183) and the
`output_function` branch (reach_blocks.jl:152-155,191-199) -- vs the oracle, plus the three check-mode known answers
the reference's own tests pin (/root/reference/test/Reachability/solve_continuous.jl:49-75).
Tolerance 1e-10 relative + 1e-12 absolute; violation indices must be identical."""
import numpy as np
import pytest

from oracle import reach_oracle as orc
from tests.util import assert_close

pytestmark = pytest.mark.gpu

A4 = np.array([[0.0509836, 0.168159, 0.95246, 0.33644],
               [0.42377, 0.67972, 0.129232, 0.126662],
               [0.518654, 0.981313, 0.489854, 0.588326],
               [0.38318, 0.616014, 0.518412, 0.778765]])


def _problem(seed, n, m, delta=0.01):
    g = np.random.default_rng(seed)
    A = g.standard_normal((n, n)) / np.sqrt(n) - 1.2 * np.eye(n)
    X0 = orc.Box(g.standard_normal(n), np.abs(g.standard_normal(n)) * 0.1)
    B = g.standard_normal((n, m)) if m else None
    U = orc.Box(g.standard_normal(m), np.abs(g.standard_normal(m)) * 0.3) if m else None
    return orc.discretize_box(A, X0, delta, B, U)


@pytest.mark.parametrize("ell,bound,rows,sizes,expected", [
    ([1., 1.], 2., [0, 1], [2], 1),                  # x1 + x2 <= 2 is violated at the first step (only block 1 computed)
    ([1., 1., 0., 0.], 2., [0, 1, 2, 3], [2, 2], 1), # same with both blocks computed
    ([1., -1., 0., 0.], 2., [0, 1, 2, 3], [2, 2], 0),  # x1 - x2 <= 2 holds  (CheckSolution.violation == -1)
])
def test_reference_known_answers(ctx, ell, bound, rows, sizes, expected):
    D = orc.discretize_box(A4, orc.Box(np.ones(4), np.full(4, 0.1)), 0.01)
    N = orc.n_steps_bffpsv18(0.1, 0.01)
    got = ctx.bffpsv18_check(D.Phi, D.c0, D.r0, N, ell, bound, rows=rows, block_sizes=sizes)
    assert got == expected
    assert got == orc.bffpsv18_check(D, N, np.array(ell), bound, rows, block_sizes=sizes)


@pytest.mark.parametrize("n,m,N,sizes", [(6, 0, 40, [2, 2, 2]), (6, 2, 40, [1] * 6), (9, 1, 70, [2, 3, 4]), (20, 3, 130, [20]),
                                         (33, 2, 90, [2] * 16 + [1]), (64, 4, 300, None)])
def test_check_first_violation_matches_oracle(ctx, n, m, N, sizes):
    D = _problem(500 + n, n, m)
    g = np.random.default_rng(n)
    rows = np.arange(n)
    ell = g.standard_normal(n)
    # pick bounds around the trajectory of rho so that some are violated early, late and never
    osizes = sizes if sizes is not None else [1] * n      # device convention: no block sizes = 1-D blocks
    C, R = orc.bffpsv18_output_map(D, N, ell[None, :], rows, osizes)
    rho = (C + R)[:, 0]
    for bound in (rho.min() - 1.0, np.quantile(rho, 0.3), np.quantile(rho, 0.9), rho.max() + 1.0, rho[0] + 1e-9):
        for eager in (True, False):
            want = orc.bffpsv18_check(D, N, ell, bound, rows, eager=eager, block_sizes=osizes)
            got = ctx.bffpsv18_check(D.Phi, D.c0, D.r0, N, ell, bound, D.MU, D.cU, D.rU, D.rpsi, rows, eager=eager,
                                     block_sizes=sizes)
            assert got == want, (bound, eager, got, want)


@pytest.mark.parametrize("n,m,N,q,sizes,rows", [(4, 0, 10, 2, [2, 2], None), (6, 2, 40, 1, [1] * 6, None), (9, 1, 70, 3, [2, 3, 4], None),
                                                (20, 3, 130, 8, [20], None), (33, 2, 90, 2, [2] * 16 + [1], None),
                                                (40, 3, 50, 2, [2, 2, 2], [4, 5, 10, 11, 38, 39]), (130, 1, 200, 2, None, None),
                                                (14, 2, 25, 11, [2] * 7, None), (10, 0, 12, 17, [5, 5], None)])    # q > 8: groups of 8 rows
def test_output_map_matches_oracle(ctx, n, m, N, q, sizes, rows):
    D = _problem(600 + n, n, m)
    g = np.random.default_rng(n + q)
    rows = np.arange(n) if rows is None else np.asarray(rows)
    M = g.standard_normal((q, len(rows)))
    M[0, :] = 0.0
    M[0, 0] = 1.0          # a coordinate projection row, as `:projection_matrix` typically has
    Co, Ro = orc.bffpsv18_output_map(D, N, M, rows, sizes if sizes is not None else [1] * len(rows))
    C, R = ctx.bffpsv18_output_map(D.Phi, D.c0, D.r0, N, M, D.MU, D.cU, D.rU, D.rpsi, rows, block_sizes=sizes)
    assert_close(C.T, Co, "centres")
    assert_close(R.T, Ro, "radii")


def test_output_map_tighter_than_boxing(ctx):
    """|M P| r0 <= |M| |P| r0: the lazy radius never exceeds the radius obtained from the boxed blocks."""
    D = _problem(7, 12, 2)
    rows = np.arange(12)
    M = np.random.default_rng(1).standard_normal((2, 12))
    C, R = ctx.bffpsv18_output_map(D.Phi, D.c0, D.r0, 50, M, D.MU, D.cU, D.rU, D.rpsi, rows, block_sizes=[12])
    Cb, Rb = ctx.bffpsv18_boxes(D.Phi, D.c0, D.r0, 50, D.MU, D.cU, D.rU, D.rpsi)
    assert_close(C, M @ Cb, "centres agree")
    assert np.all(R <= np.abs(M) @ Rb * (1 + 1e-12) + 1e-15)
