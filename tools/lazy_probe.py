"""Lazy-expm / sparse path on config C4 (n=10913 MNA5-shaped): sparse discretisation + rows of interest advanced by the
action of e^{δAᵀ}, against the dense device path on the same rows."""
import json, sys, time
from pathlib import Path
import numpy as np
import scipy.sparse as sp
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import reachability_jl_b200 as rb
from reachability_jl_b200 import workloads as wl

R = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
check_dense = int(sys.argv[3]) if len(sys.argv) > 3 else 1
w = wl.mna5()
n = w.n
rows = None if R >= n else np.unique(np.concatenate([np.arange(10), np.linspace(0, n - 1, max(R - 10, 1)).astype(int)]))[:R].astype(np.int32)
As = sp.csc_matrix(w.A)
print(f"{w.name}: n={n}, nnz={As.nnz}, rows of interest {n if rows is None else len(rows)}, N={N}", flush=True)
with rb.Context(0) as ctx:
    S = ctx.sparse(As)
    for it in range(2):
        t0 = time.time()
        Ds = S.discretize_box(w.delta, w.c, w.r, w.B, w.cU, w.rU)
        print(f"sparse discretisation {time.time() - t0:.3f} s", flush=True)
    for it in range(2):
        t0 = time.time()
        C, Rr = S.bffpsv18_boxes(w.delta, Ds["c0"], Ds["r0"], N, Ds["MU"], Ds["cU"], Ds["rU"], Ds["rpsi"], rows)
        t = S.timing()
        t.update(wall_s=time.time() - t0, steps_per_s=(N - 1) / (t["run_ms"] * 1e-3), gbs=t["bytes"] / (t["run_ms"] * 1e-3) / 1e9,
                 products_per_step=t["products"] / (N - 1))
        print(json.dumps(t), flush=True)
    if check_dense:
        t0 = time.time()
        Dd = ctx.discretize_box(w.A, w.delta, w.c, w.r, w.B, w.cU, w.rU)
        print(f"dense device discretisation {time.time() - t0:.2f} s", flush=True)
        worst = 0.0
        for key in ("c0", "r0", "rE", "rpsi"):
            u = np.max(np.abs(Ds[key] - Dd[key]) / (1e-10 * np.abs(Dd[key]) + 1e-12))
            worst = max(worst, u)
            print(f"  {key}: {u:.3e} tolerance units vs the dense device discretisation")
        Nd = min(N, check_dense if check_dense > 1 else N)
        plan = ctx.boxplan(Dd["Phi"], Dd["c0"], Dd["r0"], Nd, Dd["MU"], Dd["cU"], Dd["rU"], Dd["rpsi"], rows)
        plan.run()
        tt = plan.timing()
        Cd, Rd = plan.fetch()
        plan.close()
        uc = np.max(np.abs(C[:, :Nd] - Cd) / (1e-10 * np.abs(Cd) + 1e-12))
        ur = np.max(np.abs(Rr[:, :Nd] - Rd) / (1e-10 * np.abs(Rd) + 1e-12))
        print(f"dense device path on the same rows: {(Nd - 1) / (tt['run_ms'] * 1e-3):.1f} steps/s over {Nd} steps; lazy vs dense: centres "
              f"{uc:.3e}, radii {ur:.3e} tolerance units", flush=True)
    S.close()
