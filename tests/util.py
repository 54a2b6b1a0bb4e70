"""Helpers shared by the parity tests."""
import numpy as np

RTOL = 1e-10   # north_star tolerance: 1e-10 relative + 1e-12 absolute
ATOL = 1e-12


def tol_units(got, want):
    """max |got - want| / (RTOL |want| + ATOL): <= 1 means inside the north_star tolerance."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    if got.size == 0:
        return 0.0
    return float(np.max(np.abs(got - want) / (RTOL * np.abs(want) + ATOL)))


def assert_close(got, want, what=""):
    u = tol_units(got, want)
    assert u <= 1.0, f"{what}: {u:.3g} tolerance units (1e-10 rel + 1e-12 abs)"
