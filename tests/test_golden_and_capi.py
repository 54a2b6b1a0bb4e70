"""CPU-side checks: the oracle reproduces the committed golden fixtures (tests/golden/make_golden.py) bit for bit
up to BLAS summation order, and the C-ABI shared library loads and exports every symbol include/reach.h declares."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from oracle import reach_oracle as orc
from tests.util import assert_close

ROOT = Path(__file__).resolve().parent.parent
GOLDEN = ROOT / "tests" / "golden"


@pytest.mark.parametrize("name", ["lcs4", "clccs4", "doc5"])
def test_oracle_reproduces_golden(name):
    z = np.load(GOLDEN / f"{name}.npz")
    X0 = orc.Box(z["c"], z["r"])
    B = z["B"] if "B" in z else None
    U = orc.Box(z["cU"], z["rU"]) if "B" in z else None
    delta, T = float(z["delta"]), float(z["T"])
    for alg in ("forward", "backward"):
        D = orc.discretize_box(z["A"], X0, delta, B, U, algorithm=alg)
        assert_close(D.Phi, z[f"{alg}_Phi"], "Phi")
        assert_close(D.Phi2abs, z[f"{alg}_Phi2abs"], "Phi2abs")
        assert_close(D.c0, z[f"{alg}_c0"], "c0")
        assert_close(D.r0, z[f"{alg}_r0"], "r0")
        C, R = orc.bffpsv18_reach(D, orc.n_steps_bffpsv18(T, delta))
        assert_close(C, z[f"{alg}_centers"], "centers")
        assert_close(R, z[f"{alg}_radii"], "radii")
    Dz = orc.discretize_zonotope(z["A"], X0, delta, B, U)
    Z = orc.glgm06_reach(Dz, orc.n_steps_glgm06(T, delta), max_order=int(z["z_max_order"]))
    assert [zz.ngens for zz in Z] == list(z["z_p"])
    o = 0
    for k, zz in enumerate(Z):
        assert_close(zz.c, z["z_C"][:, k], f"centre {k + 1}")
        assert_close(zz.G, z["z_G"][:, o:o + zz.ngens], f"generators {k + 1}")
        o += zz.ngens


def _declared_symbols():
    text = (ROOT / "include" / "reach.h").read_text()
    return sorted(set(re.findall(r"REACH_API\s+[\w\s\*]+?\b(reach_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    import reachability_jl_b200 as rb
    assert sorted(rb.EXPORTED_SYMBOLS) == _declared_symbols()


def test_library_loads_and_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(str(lib_path))
    names = _declared_symbols()
    assert len(names) >= 30
    for name in names:
        assert hasattr(lib, name), f"libreach_sm100a.so does not export {name}"
    assert lib.reach_version() >= 100


def test_no_cpu_fallback_without_gpu(lib_path):
    """Without a CUDA device the context cannot be created and the error says so (no CPU path exists)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    import reachability_jl_b200 as rb
    with pytest.raises(rb.ReachError) as ei:
        rb.Context(0)
    assert ei.value.code == 2 and "no CUDA device" in str(ei.value)


def test_product_never_imports_the_oracle():
    for f in (ROOT / "reachability.jl_b200").rglob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f
