"""C2 flowpipe (400 steps) + on-device projection: timing of reach_flowpipe_project."""
import sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import reachability_jl_b200 as rb
from reachability_jl_b200 import workloads as wl
from oracle import reach_oracle as orc

N = int(sys.argv[1]) if len(sys.argv) > 1 else 400
w = wl.iss()
Phi = orc.exp_Adelta(w.A, w.delta)
P2 = orc.phi12_series(np.abs(w.A), w.delta)[2]
D = orc.discretize_zonotope(w.A, orc.Box(w.c, w.r), w.delta, w.B, orc.Box(w.cU, w.rU), Phi=Phi, Phi2abs=P2)
Om = orc.reduce_order(D.Omega0, w.max_order)
with rb.Context(0) as ctx:
    fp = ctx.glgm06(D.Phi, Om.c, Om.G, N, w.max_order, D.V.c, D.V.G, 0)
    fp.run()
    M = np.zeros((2, w.n)); M[0, 0] = 1.0; M[1, w.n // 2] = 1.0
    for it in range(3):
        t0 = time.time()
        C, R = fp.project(M)
        dt = time.time() - t0
        print(f"project {N} reach sets: {dt * 1e3:.2f} ms wall ({8.0 * w.n * fp.ngens().sum() / dt / 1e9:.0f} GB/s of generators read)", flush=True)
    lo, hi = fp.box(N)
    print("check vs box hull:", np.max(np.abs((C[:, -1] - R[:, -1]) - lo[[0, w.n // 2]])), np.max(np.abs((C[:, -1] + R[:, -1]) - hi[[0, w.n // 2]])))
    fp.close()
