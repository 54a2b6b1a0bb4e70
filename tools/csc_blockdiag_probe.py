"""Sparse twin of reach_blocks! on a block-diagonal Phi (the decoupled modes of the ISS-shaped system under BFFPSV18, all 270
1-D blocks, N = 2000): block-diagonal path against the dense DMMA recurrence, both through the host-buffer C-ABI calls."""
import sys, time
from pathlib import Path
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import reachability_jl_b200 as rb
from reachability_jl_b200 import workloads as wl

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n_modes = int(sys.argv[2]) if len(sys.argv) > 2 else 135
w = wl.iss(n_modes=n_modes)
with rb.Context(0) as ctx:
    D = ctx.discretize_box(w.A, w.delta, w.c, w.r, w.B, w.cU, w.rU)
    Phi = np.where(np.abs(D["Phi"]) > 0, D["Phi"], 0.0)
    S = sp.csc_matrix(Phi)
    print("n", w.n, "nnz(Phi)", S.nnz, "N", N)
    for name, call in (("block-diagonal path (reach_bffpsv18_boxes_csc)", lambda: ctx.bffpsv18_boxes_csc(S, D["c0"], D["r0"], N, D["MU"], D["cU"], D["rU"], D["rpsi"])),
                       ("dense DMMA recurrence (reach_bffpsv18_boxes)", lambda: ctx.bffpsv18_boxes(Phi, D["c0"], D["r0"], N, D["MU"], D["cU"], D["rU"], D["rpsi"]))):
        call()
        t0 = time.perf_counter()
        reps = 5 if w.n <= 512 else 2
        for _ in range(reps):
            C, R = call()
        dt = (time.perf_counter() - t0) / reps
        print(f"{name}: {dt * 1e3:.2f} ms per flowpipe end to end = {(N - 1) / dt:.0f} steps/s", ctx.csc_info() if "csc" in name else "")
