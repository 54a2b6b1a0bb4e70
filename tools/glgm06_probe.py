"""Timing probe of the GLGM06 device path on config C2 (n=270 ISS-shaped, max_order 10)."""
import json, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import reachability_jl_b200 as rb
from reachability_jl_b200 import workloads as wl
from oracle import reach_oracle as orc

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
check = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
w = wl.iss()
t0 = time.time()
Phi = orc.exp_Adelta(w.A, w.delta)
P2 = orc.phi12_series(np.abs(w.A), w.delta)[2]
D = orc.discretize_zonotope(w.A, orc.Box(w.c, w.r), w.delta, w.B, orc.Box(w.cU, w.rU), Phi=Phi, Phi2abs=P2)
Om = orc.reduce_order(D.Omega0, w.max_order)
print("discretize (oracle, host)", round(time.time() - t0, 2), "s; p0", Om.ngens, "pV", D.V.ngens, flush=True)
with rb.Context(0) as ctx:
    fp = ctx.glgm06(D.Phi, Om.c, Om.G, N, w.max_order, D.V.c, D.V.G, 0)
    print("generator map:", fp.map_info(), "(REACH_GLGM06_DENSE=1 forces the DMMA GEMM)", flush=True)
    for it in range(reps):
        l0 = ctx.launch_count()
        fp.run()
        t = fp.timing()
        t["launches"] = ctx.launch_count() - l0
        t["steps_per_s"] = (N - 1) / (t["run_ms"] * 1e-3)
        print(json.dumps(t), flush=True)
    p = fp.ngens()
    print("ngens first/last", p[:12], p[-3:])
    if check:
        t0 = time.time()
        R = orc.glgm06_reach_inhomog(Om, D.V, D.Phi, check, w.max_order)
        dt = time.time() - t0
        print(f"oracle {check} steps in {dt:.2f} s = {(check - 1) / dt:.1f} steps/s", flush=True)
        worst = 0.0
        for k in range(1, check + 1):
            c, G = fp.step(k)
            Zo = R[k - 1]
            assert G.shape == Zo.G.shape, (k, G.shape, Zo.G.shape)
            u = max(np.max(np.abs(c - Zo.c) / (1e-10 * np.abs(Zo.c) + 1e-12)), np.max(np.abs(G - Zo.G) / (1e-10 * np.abs(Zo.G) + 1e-12)))
            worst = max(worst, u)
            if u > 1:
                print("step", k, "tolerance units", u)
                break
        print("worst tolerance units over", check, "steps:", worst)
    fp.close()
