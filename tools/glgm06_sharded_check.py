"""2+ GPU check of the column-sharded GLGM06 flowpipe against the oracle (run under torchrun on a multi-GPU box)."""
import os, sys, time
from pathlib import Path
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import reachability_jl_b200 as rb
from reachability_jl_b200 import sharding, workloads as wl
from oracle import reach_oracle as orc

rank, world, dev = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
full = len(sys.argv) > 1 and sys.argv[1] == "c2"
cases = []
g = np.random.default_rng(5)
for (n, m, N, r, rzg) in [(13, 3, 40, 3, True), (20, 20, 30, 4, False), (33, 2, 50, 6, True), (64, 5, 70, 3, True)]:
    A = g.standard_normal((n, n)) / np.sqrt(n) - 1.2 * np.eye(n)
    X0 = orc.Box(g.standard_normal(n), np.abs(g.standard_normal(n)) * 0.1)
    B = g.standard_normal((n, m)); U = orc.Box(g.standard_normal(m), np.abs(g.standard_normal(m)) * 0.3)
    cases.append((f"random n={n}", A, X0, B, U, 0.01, N, r, rzg, N))
if full:
    w = wl.iss()
    cases.append(("C2-iss n=270", w.A, orc.Box(w.c, w.r), w.B, orc.Box(w.cU, w.rU), w.delta, 2000, 10, True, 150))
ok = True
torch_stream = torch.cuda.Stream(device=dev)          # the library borrows it, NCCL orders its collectives against it
torch.cuda.set_stream(torch_stream)
with rb.Context(dev, stream=torch_stream.cuda_stream) as ctx:
    for (name, A, X0, B, U, delta, N, r, rzg, ncheck) in cases:
        D = orc.discretize_zonotope(A, X0, delta, B, U, rzg=rzg)
        Om = orc.reduce_order(D.Omega0, r, rzg)
        fp = ctx.glgm06_sharded(D.Phi, Om.c, Om.G, N, r, D.V.c, D.V.G, rank, world, flags=0 if rzg else 1)
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        sharding.run_sharded(fp, dist)
        torch.cuda.synchronize(); dist.barrier()
        dt = time.perf_counter() - t0
        p = fp.ngens()
        R = orc.glgm06_reach_inhomog(Om, D.V, D.Phi, ncheck, r, rzg)
        worst = 0.0
        for k in sorted(set([1, 2, 3, ncheck // 2, ncheck - 1, ncheck])):
            c, G = sharding.gather_sharded_step(fp, k, int(p[k - 1]), dist, torch.device("cuda", dev))
            Zo = R[k - 1]
            if G.shape != Zo.G.shape:
                ok = False
                print(f"[rank {rank}] {name}: step {k} shape {G.shape} vs oracle {Zo.G.shape}", flush=True)
                continue
            u = max(np.max(np.abs(c - Zo.c) / (1e-10 * np.abs(Zo.c) + 1e-12)), np.max(np.abs(G - Zo.G) / (1e-10 * np.abs(Zo.G) + 1e-12)))
            worst = max(worst, float(u))
        ok = ok and worst <= 1.0
        if rank == 0:
            print(f"{name}: world={world} N={N} sharded run {dt * 1e3:.1f} ms ({(N - 1) / dt:.0f} steps/s), worst tolerance units {worst:.3g}", flush=True)
        fp.close()
if rank == 0:
    print("SHARDED CHECK", "OK" if ok else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
