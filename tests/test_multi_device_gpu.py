"""GPU: one process, one `reach_ctx` over several devices (reach_ctx_create_multi): `reach_bffpsv18_boxes` shards the rows of
interest inside the library, one host thread per device, no collective.

On a single-GPU box the same device is listed twice or three times -- the sharding, threading and strided gather are
exactly the multi-GPU code path, the shards merely share the device; with >= 2 GPUs the real thing runs too.
With the stack depth pinned (REACH_BFFPSV18_BLOCK) every device takes the same doubling path as the single-device run and
the results are bit-identical; with the per-shard heuristic they agree to rounding."""
import numpy as np
import pytest
import scipy.sparse as sp

import reachability_jl_b200 as rb
from oracle import reach_oracle as orc
from tests.util import assert_close

pytestmark = pytest.mark.gpu


def _problem(n, m, seed=5):
    g = np.random.default_rng(seed)
    A = g.standard_normal((n, n)) / np.sqrt(n) - 1.2 * np.eye(n)
    X0 = orc.Box(g.standard_normal(n), np.abs(g.standard_normal(n)) * 0.1)
    B = g.standard_normal((n, m))
    U = orc.Box(g.standard_normal(m), np.abs(g.standard_normal(m)) * 0.2)
    return orc.discretize_box(A, X0, 0.01, B, U)


def _devices():
    import torch
    lists = [[0, 0], [0, 0, 0]]
    if torch.cuda.device_count() >= 2:
        lists.append([0, 1])
    if torch.cuda.device_count() >= 4:
        lists.append([0, 1, 2, 3])
    return lists


@pytest.mark.parametrize("n,m,N,nrows", [(40, 2, 60, 40), (96, 1, 130, 50), (7, 1, 20, 3)])
def test_multi_device_boxes_equal_single_device_and_oracle(ctx, n, m, N, nrows, monkeypatch):
    D = _problem(n, m)
    rows = np.sort(np.random.default_rng(n).choice(n, nrows, replace=False)).astype(np.int32)
    Co, Ro = orc.bffpsv18_reach(D, N, rows)
    monkeypatch.setenv("REACH_BFFPSV18_BLOCK", "8")
    C1, R1 = ctx.bffpsv18_boxes(D.Phi, D.c0, D.r0, N, D.MU, D.cU, D.rU, D.rpsi, rows)
    for devs in _devices():
        with rb.Context(devs) as mctx:
            assert mctx.device_count() == len(devs)
            l0 = mctx.launch_count()
            Cm, Rm = mctx.bffpsv18_boxes(D.Phi, D.c0, D.r0, N, D.MU, D.cU, D.rU, D.rpsi, rows)
            assert mctx.launch_count() > l0
            assert np.array_equal(Cm, C1) and np.array_equal(Rm, R1), f"devices {devs}: not bit-identical to one device"
            # the primary device of the handle still serves every other entry point
            assert_close(mctx.expm(np.diag([-1.0, -2.0]), 0.5), np.diag(np.exp([-0.5, -1.0])), "expm on the primary device")
    monkeypatch.delenv("REACH_BFFPSV18_BLOCK")
    with rb.Context([0, 0]) as mctx:               # per-shard stack depth: same sets to rounding
        Cm, Rm = mctx.bffpsv18_boxes(D.Phi, D.c0, D.r0, N, D.MU, D.cU, D.rU, D.rpsi, rows)
    assert_close(Cm.T, Co, "centres vs oracle")
    assert_close(Rm.T, Ro, "radii vs oracle")


def test_multi_device_csc_twin_and_homogeneous(ctx):
    D = _problem(30, 1, seed=9)
    Phi = D.Phi.copy()
    Phi[np.abs(Phi) < 0.02] = 0.0
    D.Phi = Phi
    N = 25
    with rb.Context([0, 0, 0]) as mctx:
        C1, R1 = ctx.bffpsv18_boxes_csc(sp.csc_matrix(Phi), D.c0, D.r0, N)
        Cm, Rm = mctx.bffpsv18_boxes_csc(sp.csc_matrix(Phi), D.c0, D.r0, N)
    assert_close(Cm, C1, "csc, homogeneous: centres")
    assert_close(Rm, R1, "csc, homogeneous: radii")


def test_multi_device_context_reports_errors_of_a_shard(ctx):
    with pytest.raises(rb.ReachError):
        rb.Context([0, 4096])                      # a device that does not exist: creation fails as a whole
    D = _problem(10, 1)
    with rb.Context([0, 0]) as mctx:
        with pytest.raises(rb.ReachError) as ei:
            mctx.bffpsv18_boxes(D.Phi, D.c0, D.r0, 5, D.MU, D.cU, D.rU, D.rpsi, np.array([0, 1, 99], dtype=np.int32))
        assert "row index out of range" in str(ei.value)
