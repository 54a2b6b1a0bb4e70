"""ctypes binding of libreach_sm100a.so (include/reach.h).  There is no CPU fallback: if the shared library
is missing or no B200 is visible, every entry point raises."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

LIB_PATH = Path(__file__).resolve().parent / "libreach_sm100a.so"

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
i64, i32, f64, vp = C.c_int64, C.c_int32, C.c_double, C.c_void_p


class ReachError(RuntimeError):
    """Non-zero status from libreach_sm100a (the Julia shim turns these into `error(reach_last_error())`)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"libreach_sm100a status {code}: {msg}")
        self.code = code


# name -> argtypes (restype is int32 unless listed in _RESTYPES)
_MATIN = [c_double_p, i64]
_BOX_ARGS = [vp, i64, c_double_p, i64, c_double_p, c_double_p, i64, c_double_p, i64, c_double_p, c_double_p, c_double_p,
             i64, c_int32_p, i64]
_SIGNATURES = {
    "reach_version": [],
    "reach_ctx_create": [i32, C.POINTER(vp)],
    "reach_ctx_create_on_stream": [i32, vp, C.POINTER(vp)],
    "reach_ctx_create_multi": [c_int32_p, i32, C.POINTER(vp)],
    "reach_ctx_device_count": [vp, c_int32_p],
    "reach_ctx_destroy": [vp],
    "reach_last_error": [vp],
    "reach_sync": [vp],
    "reach_launch_count": [vp, C.POINTER(C.c_uint64)],
    "reach_expm": [vp, i64, c_double_p, i64, f64, c_double_p, i64],
    "reach_phi12": [vp, i64, c_double_p, i64, f64, c_double_p, i64, c_double_p, i64],
    "reach_discretize_box": [vp, i64, c_double_p, i64, f64, i32, c_double_p, c_double_p, i64, c_double_p, i64,
                             c_double_p, c_double_p, c_double_p, i64, c_double_p, i64, c_double_p, c_double_p,
                             c_double_p, c_double_p, c_double_p],
    "reach_discretize_zonotope": [vp, i64, c_double_p, i64, f64, i32, c_double_p, c_double_p, i64, c_double_p, i64, c_double_p,
                                  c_double_p, i32, c_double_p, i64, c_double_p, c_double_p, i64, c_int64_p, c_double_p, c_double_p,
                                  i64, c_int64_p],
    "reach_discretize_inputs_box": [vp, i64, c_double_p, i64, f64, i64, c_double_p, i64, i64, c_double_p, i64, c_double_p, i64,
                                    c_double_p, i64, c_double_p],
    "reach_bffpsv18_boxes": _BOX_ARGS + [c_double_p, c_double_p, c_int64_p],
    "reach_bffpsv18_boxes_csc": [vp, i64, i64, c_int64_p, c_int64_p, c_double_p, c_double_p, c_double_p, i64, c_double_p, i64,
                                 c_double_p, c_double_p, c_double_p, i64, c_int32_p, i64, c_double_p, c_double_p, c_int64_p],
    "reach_bffpsv18_csc_info": [vp, c_int64_p, c_int32_p],
    "reach_boxplan_create": _BOX_ARGS + [C.POINTER(vp)],
    "reach_boxplan_run": [vp],
    "reach_boxplan_fetch": [vp, c_double_p, c_double_p],
    "reach_boxplan_device_results": [vp, C.POINTER(c_double_p), C.POINTER(c_double_p)],
    "reach_boxplan_timing": [vp, c_double_p, c_double_p, c_int64_p, c_double_p],
    "reach_boxplan_destroy": [vp],
    "reach_bffpsv18_check": _BOX_ARGS + [i64, c_int32_p, c_double_p, f64, i32, c_int64_p],
    "reach_bffpsv18_output_map": _BOX_ARGS + [i64, c_int32_p, i64, c_double_p, i64, c_double_p, c_double_p],
    "reach_sparse_create": [vp, i64, i64, c_int64_p, c_int64_p, c_double_p, C.POINTER(vp)],
    "reach_sparse_free": [vp],
    "reach_expmv": [vp, f64, i32, i64, c_double_p, i64, c_double_p, i64],
    "reach_discretize_box_sparse": [vp, f64, i32, c_double_p, c_double_p, i64, c_double_p, i64, c_double_p, c_double_p,
                                    c_double_p, c_double_p, c_double_p, c_double_p, c_double_p],
    "reach_bffpsv18_boxes_lazy": [vp, f64, c_double_p, c_double_p, i64, c_double_p, i64, c_double_p, c_double_p, c_double_p,
                                  i64, c_int32_p, i64, c_double_p, c_double_p, c_int64_p],
    "reach_bffpsv18_output_map_lazy": [vp, f64, c_double_p, c_double_p, i64, c_double_p, i64, c_double_p, c_double_p, c_double_p,
                                       i64, c_int32_p, i64, i64, c_int32_p, i64, c_double_p, i64, c_double_p, c_double_p],
    "reach_bffpsv18_check_lazy": [vp, f64, c_double_p, c_double_p, i64, c_double_p, i64, c_double_p, c_double_p, c_double_p,
                                  i64, c_int32_p, i64, i64, c_int32_p, c_double_p, f64, i32, c_int64_p],
    "reach_sparse_timing": [vp, c_double_p, c_int64_p, c_double_p],
    "reach_glgm06_create": [vp, i64, c_double_p, i64, i64, c_double_p, c_double_p, i64, i64, c_double_p, c_double_p,
                            i64, i64, i64, i32, C.POINTER(vp)],
    "reach_glgm06_create_sharded": [vp, i64, c_double_p, i64, i64, c_double_p, c_double_p, i64, i64, c_double_p, c_double_p,
                                    i64, i64, i64, i32, i32, i32, C.POINTER(vp)],
    "reach_glgm06_shard_nbatches": [vp, c_int64_p],
    "reach_glgm06_shard_phase": [vp, i64, i32],
    "reach_glgm06_shard_phase_on": [vp, i64, i32, vp],
    "reach_glgm06_shard_buffer": [vp, i64, i32, C.POINTER(vp), c_int64_p, c_int32_p],
    "reach_glgm06_run": [vp],
    "reach_glgm06_run_fetch": [vp, c_double_p, c_double_p, i64],
    "reach_flowpipe_ngens": [vp, c_int32_p],
    "reach_flowpipe_step": [vp, i64, c_double_p, c_double_p, i64],
    "reach_flowpipe_fetch": [vp, c_double_p, c_double_p, i64],
    "reach_flowpipe_box": [vp, i64, c_double_p, c_double_p],
    "reach_flowpipe_project": [vp, i64, c_double_p, i64, c_double_p, c_double_p, c_double_p, i64],
    "reach_flowpipe_selection": [vp, i64, i32, c_int32_p, c_int32_p, c_int32_p, c_double_p],
    "reach_flowpipe_timing": [vp, c_double_p, c_double_p, c_int64_p, c_double_p, c_double_p, c_double_p],
    "reach_flowpipe_phase_ms": [vp, c_double_p, c_double_p],
    "reach_flowpipe_map_info": [vp, c_int32_p, c_int32_p],
    "reach_flowpipe_chain_info": [vp, c_int64_p, c_int64_p],
    "reach_flowpipe_set_serial": [vp, i32],
    "reach_flowpipe_free": [vp],
    "reach_reduce_order": [vp, i64, i64, c_double_p, c_double_p, i64, i64, i32, c_double_p, i64, c_int64_p, c_int32_p],
    "reach_ensemble_create": [vp, i64, c_double_p, i64, i64, i64, c_double_p, c_double_p, i64, C.POINTER(vp)],
    "reach_ensemble_run": [vp],
    "reach_ensemble_step": [vp, i64, i64, c_double_p, c_double_p],
    "reach_ensemble_boxes": [vp, i64, c_double_p, c_double_p],
    "reach_ensemble_boxes_all": [vp, c_double_p, c_double_p],
    "reach_ensemble_timing": [vp, c_double_p, c_double_p],
    "reach_ensemble_free": [vp],
    "reach_dgemm_left": [vp, i64, i64, i64, c_double_p, i64, c_double_p, i64, c_double_p, i64, i32, c_double_p],
    "reach_dgemm": [vp, i64, i64, i64, c_double_p, i64, c_double_p, i64, c_double_p, i64, i32, c_double_p],
}
_RESTYPES = {
    "reach_ctx_destroy": None, "reach_boxplan_destroy": None, "reach_flowpipe_free": None,
    "reach_ensemble_free": None, "reach_last_error": C.c_char_p, "reach_sparse_free": None,
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)
_lib = None


def load() -> C.CDLL:
    """Load the shared library (built in-tree by ``__graft_entry__.build()``); raise if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ReachError(2, f"{LIB_PATH} not found: build it with `python __graft_entry__.py build` "
                            "(libreach_sm100a has no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int32)
    _lib = lib
    return lib


def dptr(a):
    """Pointer to a float64 numpy array (or NULL for None)."""
    if a is None:
        return None
    assert a.dtype == np.float64
    return a.ctypes.data_as(c_double_p)


def iptr(a):
    if a is None:
        return None
    assert a.dtype == np.int32
    return a.ctypes.data_as(c_int32_p)


def fmat(a) -> np.ndarray:
    """Column-major float64 copy/view (Julia layout)."""
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def fvec(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel())
