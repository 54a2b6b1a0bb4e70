"""Per-call wall time of create + run + project on C2 (which part of the projected end-to-end number jitters?)."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import reachability_jl_b200 as rb
from reachability_jl_b200 import api, workloads as wl
w = wl.iss()
N = 2000
with rb.Context(0) as ctx:
    Phi, (c0, G0), V = api.discretize_zonotope(ctx, w.A, w.delta, w.c, w.r, w.B, (w.cU, w.rU))
    G0red, _ = ctx.reduce_order(c0, G0, w.max_order)
    P2 = np.zeros((2, w.n)); P2[0, 0] = 1.0; P2[1, w.n // 2] = 1.0
    for i in range(12):
        t0 = time.perf_counter()
        f = ctx.glgm06(Phi, c0, G0red, N, w.max_order, V[0], V[1], 0)
        t1 = time.perf_counter()
        f.run()
        t2 = time.perf_counter()
        out = f.project(P2)
        t3 = time.perf_counter()
        f.close()
        t4 = time.perf_counter()
        print(f"call {i}: create {1e3*(t1-t0):7.2f}  run {1e3*(t2-t1):7.2f}  project {1e3*(t3-t2):7.2f}  close {1e3*(t4-t3):7.2f} ms", flush=True)
