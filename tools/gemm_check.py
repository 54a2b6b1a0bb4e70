"""GPU check of the library's DMMA GEMM against numpy (run under gpurun)."""
import ctypes, sys, time, json
import numpy as np
lib = ctypes.CDLL("reachability.jl_b200/libreach_sm100a.so")
ctx = ctypes.c_void_p()
lib.reach_last_error.restype = ctypes.c_char_p
rc = lib.reach_ctx_create(0, ctypes.byref(ctx))
assert rc == 0, lib.reach_last_error(None)
P = ctypes.POINTER(ctypes.c_double)
rng = np.random.default_rng(0)
shapes = [(8, 8, 4), (128, 128, 16), (130, 70, 33), (270, 273, 270), (1006, 1006, 1006), (48, 48, 48),
          (8640, 270, 270), (16096, 1006, 1006), (4096, 4096, 4096), (540000 // 4, 1085, 270)]
for (M, N, K) in shapes:
    A = np.asfortranarray(rng.standard_normal((M, K)))
    B = np.asfortranarray(rng.standard_normal((K, N)))
    C = np.zeros((M, N), order="F")
    ms = ctypes.c_double()
    rc = lib.reach_dgemm(ctx, ctypes.c_int64(M), ctypes.c_int64(N), ctypes.c_int64(K), A.ctypes.data_as(P), ctypes.c_int64(M),
                         B.ctypes.data_as(P), ctypes.c_int64(K), C.ctypes.data_as(P), ctypes.c_int64(M), 10, ctypes.byref(ms))
    assert rc == 0, lib.reach_last_error(ctx)
    ref = A @ B
    err = np.max(np.abs(C - ref)) / max(1.0, np.max(np.abs(ref)))
    print(json.dumps({"M": M, "N": N, "K": K, "rel_err": float(err), "ms": ms.value,
                      "tflops": 2.0 * M * N * K / (ms.value * 1e-3) / 1e12 if ms.value > 0 else None}), flush=True)
lib.reach_ctx_destroy(ctx)
