"""Timing probe of the BFFPSV18 device path on C1 (n=48) / C4 (n=10913, truncated horizon) / C3."""
import json, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import reachability_jl_b200 as rb
from reachability_jl_b200 import workloads as wl

name = sys.argv[1] if len(sys.argv) > 1 else "C4"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 0
w = {"C1": wl.building, "C3": wl.fom, "C4": wl.mna5}[name]()
if not N:
    N = int(np.ceil(w.T / w.delta))
with rb.Context(0) as ctx:
    t0 = time.time()
    D = ctx.discretize_box(w.A, w.delta, w.c, w.r, w.B, w.cU, w.rU)
    print(f"{w.name}: n={w.n} device discretisation {time.time() - t0:.2f} s (incl. H2D/D2H of the n x n matrices)", flush=True)
    plan = ctx.boxplan(D["Phi"], D["c0"], D["r0"], N, D["MU"], D["cU"], D["rU"], D["rpsi"])
    for it in range(3):
        plan.run()
        t = plan.timing()
        t["steps_per_s"] = (N - 1) / (t["run_ms"] * 1e-3)
        t["gemm_tflops"] = t["gemm_flops"] / (t["gemm_ms"] * 1e-3) / 1e12 if t["gemm_ms"] else None
        t["N"] = N
        print(json.dumps(t), flush=True)
    plan.close()
