"""compute-sanitizer target: the pipelined GLGM06 device path on a truncated C2 (n=270 ISS-shaped, max_order 10).

  compute-sanitizer --tool memcheck  python tools/sanitize_c2.py 400
  compute-sanitizer --tool racecheck python tools/sanitize_c2.py 400

Five streams, three rotating sets of batch-local arrays, mbarrier rings and bulk copies are all exercised (400 steps =
batches of 1, 2, 4, ..., 64, 64, 64, 64, 64 steps).  Inputs come from the device discretisation (no oracle here)."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import reachability_jl_b200 as rb
from reachability_jl_b200 import api, workloads as wl

N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
which = sys.argv[2] if len(sys.argv) > 2 else "glgm06"
with rb.Context(0) as ctx:
    if which == "glgm06":
        w = wl.iss()
        Phi, (c0, G0), V = api.discretize_zonotope(ctx, w.A, w.delta, w.c, w.r, w.B, (w.cU, w.rU))
        G0r, _ = ctx.reduce_order(c0, G0, w.max_order)
        fp = ctx.glgm06(Phi, c0, G0r, N, w.max_order, V[0], V[1], 0).run()
        p = fp.ngens()
        c, G = fp.step(N)
        print("GLGM06 C2 N =", N, "generators of the last set:", int(p[-1]), "finite:", bool(np.all(np.isfinite(G))),
              "launches:", ctx.launch_count())
        fp.close()
    else:
        w = wl.fom()
        D = ctx.discretize_box(w.A, w.delta, w.c, w.r, w.B, w.cU, w.rU)
        C, R = ctx.bffpsv18_boxes(D["Phi"], D["c0"], D["r0"], N, D["MU"], D["cU"], D["rU"], D["rpsi"])
        print("BFFPSV18 C3 N =", N, "finite:", bool(np.all(np.isfinite(C)) and np.all(np.isfinite(R))), "launches:", ctx.launch_count())
