"""GPU check: BFFPSV18 device path vs the numpy oracle (run under gpurun)."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
import reachability_jl_b200 as rb
from reachability_jl_b200 import workloads as wl
from oracle import reach_oracle as orc

ctx = rb.Context(0)
def run(w, N, rows=None, tag=""):
    X0 = orc.Box(w.c, w.r)
    U = orc.Box(w.cU, w.rU) if w.B is not None else None
    t0 = time.time()
    Phi = orc.exp_Adelta(w.A, w.delta)
    P2 = orc.phi12_series(np.abs(w.A), w.delta)[2]
    D = orc.discretize_box(w.A, X0, w.delta, w.B, U, Phi=Phi, Phi2abs=P2)
    t1 = time.time()
    Co, Ro = orc.bffpsv18_reach(D, N, rows)
    t2 = time.time()
    plan = ctx.boxplan(D.Phi, D.c0, D.r0, N, D.MU, D.cU, D.rU, D.rpsi, rows)
    plan.run(); plan.run()
    tm = plan.timing()
    Cg, Rg = plan.fetch()
    errc = np.max(np.abs(Cg.T - Co) / (1e-10 * np.abs(Co) + 1e-12))
    errr = np.max(np.abs(Rg.T - Ro) / (1e-10 * np.abs(Ro) + 1e-12))
    print(json.dumps(dict(tag=tag, n=w.n, N=N, nrows=Cg.shape[0], tol_units_c=float(errc), tol_units_r=float(errr),
                          oracle_disc_s=t1 - t0, oracle_loop_s=t2 - t1, gpu=tm,
                          steps_per_s=(N - 1) / (tm["run_ms"] * 1e-3),
                          gemm_tflops=tm["gemm_flops"] / (tm["gemm_ms"] * 1e-3) / 1e12 if tm["gemm_ms"] else None)), flush=True)
    plan.close()

run(wl.building(), 40, tag="C1 N=40")
run(wl.building(), 6667, tag="C1 full")
run(wl.building(), 300, rows=[0, 3, 24, 25, 47], tag="C1 rows subset")
w = wl.fom(n_pairs=20); run(w, 100, tag="fom46")
w = wl.fom(); run(w, 70, tag="C3 N=70")
run(w, 70, rows=list(range(100, 226)), tag="C3 N=70 126 rows")
# homogeneous
w2 = wl.fom(n_pairs=30); w2.B = None; run(w2, 50, tag="fom66 homog")
