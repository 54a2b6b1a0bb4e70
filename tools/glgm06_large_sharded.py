"""Large n*p GLGM06 (dense Phi, n = 1024 by default): one GPU through the pipelined engine against the generator-column-sharded
engine on WORLD_SIZE GPUs (run under torchrun; world 1 runs both engines on one GPU).
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/glgm06_large_sharded.py [n_modes] [N] [order]"""
import json, os, sys, time
from pathlib import Path
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import reachability_jl_b200 as rb
from reachability_jl_b200 import api, sharding, workloads as wl

rank, world, dev = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
n_modes = int(sys.argv[1]) if len(sys.argv) > 1 else 512
N = int(sys.argv[2]) if len(sys.argv) > 2 else 200
order = int(sys.argv[3]) if len(sys.argv) > 3 else 10
w = wl.iss_mixed(n_modes=n_modes, max_order=order)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
out = {"n": w.n, "N": N, "order": order, "world": world}
with rb.Context(dev, stream=stream.cuda_stream) as ctx:
    t0 = time.perf_counter()
    Phi, (c0, G0), V = api.discretize_zonotope(ctx, w.A, w.delta, w.c, w.r, w.B, (w.cU, w.rU))
    G0red, _ = ctx.reduce_order(c0, G0, order)
    out["discretize_s"] = time.perf_counter() - t0
    out["p0"], out["pV"] = int(G0red.shape[1]), int(V[1].shape[1])
    if rank == 0:
        fp = ctx.glgm06(Phi, c0, G0red, N, order, V[0], V[1], 0)
        best = 1e30
        for _ in range(3):
            fp.run()
            t = fp.timing()
            best = min(best, t["run_ms"])
        out["single_gpu"] = {"ms": best, "steps_per_s": (N - 1) / (best * 1e-3), "gemm_ms": t["gemm_ms"], "numerics_ms": t["numerics_ms"],
                             "chain_ms": t["chain_ms"], "gemm_tflops": t["gemm_flops"] / (best * 1e-3) * 1e-12, "map": fp.map_info()}
        out["single_gpu"]["chain_info"] = fp.chain_info()
        fp.set_serial(True); fp.run(); ts = fp.timing(); fp.set_serial(False)
        out["single_gpu"]["alone"] = {k: ts[k] for k in ("run_ms", "gemm_ms", "numerics_ms", "chain_ms")}
        pmax = int(fp.ngens().max())
        check_steps = sorted(set([1, 2, 3, 17, 64, 65, 66, 129, N // 2, N - 1, N]) & set(range(1, N + 1)))
        ref = {k: fp.step(k, int(fp.ngens()[k - 1])) for k in check_steps}
        ref_p = fp.ngens().copy()
        fp.close()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

    class _Solo:            # world 1: the sharded engine with one rank (no collectives)
        @staticmethod
        def all_reduce(t, op=None, async_op=False):
            return None
        class ReduceOp:
            SUM = None
    d = dist if world > 1 else _Solo
    fs = ctx.glgm06_sharded(Phi, c0, G0red, N, order, V[0], V[1], rank, world)
    gs = None if os.environ.get("SERIAL_PHASES") else torch.cuda.Stream(device=dev)
    sharding.run_sharded(fs, d, gemm_stream=gs)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record(stream)
    for _ in range(reps):
        sharding.run_sharded(fs, d, gemm_stream=gs)
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    out["sharded"] = {"ms": ms, "steps_per_s": (N - 1) / (ms * 1e-3), "ranks": world, "pipelined": gs is not None}
    # reach sets of the sharded run (after several repeated runs) against the single-GPU engine: same library, identical
    # selections expected -- a race in the pipelined phase order would show up as a different generator count or value
    p = fs.ngens()
    steps = [1, 2, 3, 17, 64, 65, 66, 129, N // 2, N - 1, N]
    steps = sorted(set(steps) & set(range(1, N + 1)))
    worst = 0.0
    for k in steps:
        if world > 1:
            c, G = sharding.gather_sharded_step(fs, k, int(p[k - 1]), dist, torch.device("cuda", dev))
        else:
            c, G = fs.step(k, int(p[k - 1]))
        if rank == 0:
            rc, rG = ref[k]
            u = max(np.max(np.abs(c - rc) / (1e-10 * np.abs(rc) + 1e-12)), np.max(np.abs(G - rG) / (1e-10 * np.abs(rG) + 1e-12))) if G.shape == rG.shape else float("inf")
            worst = max(worst, float(u))
    if rank == 0:
        out["generator_counts_equal"] = bool(np.array_equal(p, ref_p))
        out["steps_compared"] = steps
        out["last_step_vs_single_gpu_tolerance_units"] = worst
        out["speedup_vs_single_gpu"] = out["sharded"]["steps_per_s"] / out["single_gpu"]["steps_per_s"]
        print(json.dumps(out), flush=True)
    fs.close()
if world > 1:
    dist.destroy_process_group()
