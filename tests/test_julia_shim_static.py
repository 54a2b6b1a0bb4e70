"""CPU, static: julia/ReachSM100a.jl cannot be executed here (no Julia toolchain), so its `ccall` signatures are checked
textually against the ctypes table that the parity tests drive (reachability.jl_b200/_lib.py), and the methods it adds are
checked against the reference's own positional signatures (reach_blocks.jl:135-150, check_blocks.jl:105-116) -- the round-1
shim had a 10-argument `reach_blocks!` that the registry call `available_algorithms[...]["func"](args...)` (15 positional
arguments, reach.jl:191-193) could never dispatch to."""
import ctypes as C
import re
from pathlib import Path

from reachability_jl_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]
SHIM = (ROOT / "julia" / "ReachSM100a.jl").read_text()

_SCALARS = {"Int64": C.c_int64, "Int32": C.c_int32, "Float64": C.c_double}


def _is_pointer(ct):
    return ct is C.c_void_p or ct is C.c_char_p or (isinstance(ct, type) and issubclass(ct, C._Pointer))


def _split_types(txt):
    return [t.strip() for t in txt.replace("\n", " ").split(",") if t.strip()]


def test_every_ccall_matches_the_ctypes_signature():
    calls = re.findall(r"ccall\(\(:(\w+), LIB\),\s*(\w+),\s*\((.*?)\),\s*\n?\s*[\w.\[\]]", SHIM, flags=re.S)
    assert len(calls) >= 20
    seen = set()
    for name, ret, types in calls:
        assert name in _lib._SIGNATURES, f"{name} is not declared in the ctypes table / header"
        seen.add(name)
        want = _lib._SIGNATURES[name]
        got = _split_types(types)
        assert len(got) == len(want), f"{name}: the shim passes {len(got)} arguments, the ABI takes {len(want)}"
        for i, (jt, ct) in enumerate(zip(got, want)):
            if jt in _SCALARS:
                assert ct is _SCALARS[jt], f"{name} argument {i + 1}: Julia {jt} vs ABI {ct}"
            else:
                assert jt.startswith(("Ptr{", "Ref{")) or jt == "Cstring", f"{name} argument {i + 1}: unknown Julia type {jt}"
                assert _is_pointer(ct), f"{name} argument {i + 1}: Julia {jt} vs ABI {ct}"
        restype = _lib._RESTYPES.get(name, C.c_int32)
        assert {"Int32": C.c_int32, "Cvoid": None, "Cstring": C.c_char_p}[ret] is restype, f"{name}: return type"
    for needed in ("reach_ctx_create_multi", "reach_expm", "reach_phi12", "reach_discretize_zonotope", "reach_bffpsv18_boxes",
                   "reach_bffpsv18_output_map", "reach_bffpsv18_check", "reach_glgm06_create", "reach_glgm06_run_fetch",
                   "reach_reduce_order", "reach_flowpipe_project", "reach_sparse_create", "reach_bffpsv18_boxes_lazy",
                   "reach_bffpsv18_boxes_csc"):
        assert needed in seen, f"the shim no longer binds {needed}"


def _positional_names(src, fname, start_pat):
    """Argument names of the first method of `fname` whose text matches start_pat."""
    m = re.search(start_pat, src, flags=re.S)
    assert m, f"no method of {fname} found"
    depth, i = 0, src.index("(", m.start())    # the opening parenthesis of the argument list
    j = i
    while True:
        ch = src[j]
        depth += ch == "("
        depth -= ch == ")"
        if depth == 0:
            break
        j += 1
    args, cur, d = [], "", 0
    for ch in src[i + 1:j]:
        if ch in "({[":
            d += 1
        if ch in ")}]":
            d -= 1
        if ch == "," and d == 0:
            args.append(cur)
            cur = ""
        else:
            cur += ch
    args.append(cur)
    return [a.strip().split("::")[0].strip() for a in args if a.strip()]


def test_reach_blocks_method_has_the_reference_signature():
    ref = Path("/root/reference/src/ReachSets/ContinuousPost/BFFPSV18/reach_blocks.jl")
    want = ["ϕ", "Xhat0", "U", "overapproximate", "overapproximate_inputs", "n", "N", "output_function", "blocks", "partition",
            "dimensions", "δ", "termination", "progress_meter", "res"]              # reach_blocks.jl:135-150
    if ref.exists():                                                                 # (not on the GPU box)
        assert _positional_names(ref.read_text(), "reach_blocks!", r"# dense\nfunction reach_blocks!\(") == want
    got = _positional_names(SHIM, "reach_blocks!", r"function reach_blocks!\(ϕ::Matrix\{Float64\},")
    assert got == want and len(got) == 15


def test_check_blocks_method_has_the_reference_signature():
    ref = Path("/root/reference/src/ReachSets/ContinuousPost/BFFPSV18/check_blocks.jl")
    want = ["ϕ", "Xhat0", "U", "overapproximate_inputs", "n", "N", "blocks", "partition", "eager_checking", "prop",
            "progress_meter"]                                                        # check_blocks.jl:105-116
    if ref.exists():
        assert _positional_names(ref.read_text(), "check_blocks", r"# dense\nfunction check_blocks\(") == want
    assert _positional_names(SHIM, "check_blocks", r"function check_blocks\(ϕ::Matrix\{Float64\},") == want


def test_glgm06_methods_have_the_reference_signatures():
    for fname, want in (("reach_homog!", ["R", "Ω0", "Φ", "N", "δ"]),                                  # GLGM06/reach.jl:5-9
                        ("reach_inhomog!", ["R", "Ω0", "U", "Φ", "N", "δ", "max_order"])):             # :30-36
        pat = (r"function " if fname == "reach_inhomog!" else r"") + re.escape(fname) + r"\(R::Vector"
        assert _positional_names(SHIM, fname, pat) == want
