"""Timing probe of the batched homogeneous GLGM06 ensemble (config C5: heli28-shaped, n=28, p=85 generators)."""
import json, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import reachability_jl_b200 as rb
from reachability_jl_b200 import workloads as wl
from oracle import reach_oracle as orc

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
w = wl.heli28()
n = w.n
Phi = orc.exp_Adelta(w.A, w.delta)
cs, rs = wl.split_box(w.c, w.r, [4, 4, 4, 4, 2, 2, 2, 2])      # 4096 sub-boxes
cs, rs = cs[:, :batch], rs[:, :batch]
D0 = orc.discretize_zonotope(w.A, orc.Box(cs[:, 0], rs[:, 0]), w.delta, rzg=False)
p = D0.Omega0.ngens
c0s = np.zeros((n, batch), order="F")
G0s = np.zeros((n, p, batch), order="F")
P2 = D0.Omega0  # first member; the others share Phi / Phi2 and differ by the box
Phi2abs = orc.Phi2(np.abs(w.A), w.delta)
for b in range(batch):
    Db = orc.discretize_zonotope(w.A, orc.Box(cs[:, b], rs[:, b]), w.delta, rzg=False, Phi=Phi, Phi2abs=Phi2abs)
    c0s[:, b] = Db.Omega0.c
    G0s[:, :, b] = Db.Omega0.G
print(f"C5 ensemble: n={n}, p={p}, batch={batch}, N={N}", flush=True)
with rb.Context(0) as ctx:
    ens = ctx.ensemble(Phi, c0s, G0s, N)
    for it in range(3):
        ens.run()
        t = ens.timing()
        t["steps_per_s"] = (N - 1) / (t["run_ms"] * 1e-3)
        t["member_steps_per_s"] = t["steps_per_s"] * batch
        t["hbm_gbs"] = t["bytes_per_step"] * (N - 1) / (t["run_ms"] * 1e-3) / 1e9
        print(json.dumps(t), flush=True)
    # spot check member 3 against the oracle
    R = orc.glgm06_reach_homog(orc.Zono(c0s[:, 3], G0s[:, :, 3]), Phi, N, rzg=False)
    for k in (1, 2, N // 2, N):
        c, G = ens.step(3, k)
        u = max(np.max(np.abs(c - R[k - 1].c) / (1e-10 * np.abs(R[k - 1].c) + 1e-12)), np.max(np.abs(G - R[k - 1].G) / (1e-10 * np.abs(R[k - 1].G) + 1e-12)))
        print("member 3 step", k, "tolerance units", float(u))
    ens.close()
