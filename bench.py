#!/usr/bin/env python
"""bench.py -- flowpipe steps/s of the linear-flowpipe hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload C3|C2|C1] [--impl ours|reference]

One bench "step" = one complete flowpipe of the workload (all N_flow reach sets) computed from already
discretised inputs; `value` = flowpipe steps per second = K * (N_flow - 1) / t with the inputs resident in HBM;
`e2e` = the same through the host-buffer C-ABI call (`reach_bffpsv18_boxes` / `reach_glgm06_*`), pinned host
inputs in, host results out, every step.  N > 1 (under torchrun): BFFPSV18 shards the rows of interest across
ranks with Φ replicated and all-gathers the boxes at the end (strong scaling).

The CPU arm (`--impl reference`, and the `cpu_baseline` object of the default arm) times the oracle port of the
reference's loop (oracle/reach_oracle.py; numpy + multithreaded BLAS on all host cores) on a bounded prefix of the
same workload: the reference itself is pure Julia and no Julia toolchain exists on the GPU box.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

if "reference" in sys.argv or os.environ.get("WORLD_SIZE", "1") == "1":
    # The CPU legs use every host core: torchrun exports OMP_NUM_THREADS=1 to its workers, which would silently turn the
    # reference arm into a single-thread baseline.  Must happen before numpy / BLAS load.  (The GPU arm under torchrun keeps
    # torchrun's setting: its host side does no BLAS work inside any timed region.)
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        if "reference" in sys.argv or _v not in os.environ:
            os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# dram bytes (read + write) of the HBM-side GLGM06 kernels per flowpipe step at C2, from ncu --set full captures of one 64-step batch
# (round 2: reach sets are generator references, nothing is copied into dense slots any more):
# ell_left_score 335.2 MB + apply_r 72.0 MB + wbox_partial 38.2 MB + rank_r 3.8 + finalize_r 1.1 + sort / chain / wseq 2.0 = 452 MB / 64 steps
C2_REAL_BYTES_PER_STEP = 452.3e6 / 64
C2_REAL_BYTES_SOURCE = "profiles/r02/ncu_c2_summary.txt"
FP64_DMMA_PEAK_TFLOPS = 37.06   # measured on this pool's B200: tools/fp64_peak.cu -> profiles/r01/fp64_peak_b200.jsonl
METRIC = "flowpipe steps/s (n=1006 BFFPSV18, n=270 GLGM06) at 1/2/4/8 B200 vs CPU"


def _peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:       # NVML missing: clocks are reported as unknown
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop_evt.wait(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def pinned(shape, dtype=np.float64, order="F"):
    """numpy array over page-locked host memory (torch owns the allocation)."""
    import torch
    n = int(np.prod(shape)) if len(shape) else 1
    t = torch.empty(max(n, 1), dtype={np.float64: torch.float64, np.int32: torch.int32}[dtype], pin_memory=True)
    a = t.numpy()[:n].reshape(shape, order=order)
    a.flags.writeable = True
    _KEEP.append(t)
    return a


_KEEP = []


# ---------------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------------

def build_box_workload(name):
    """Discretised BFFPSV18 problem of config C1/C3 on the HOST, for the CPU legs only (`cpu_baseline`, `--impl reference`):
    the GPU arm prepares its inputs with the library's own device discretisation."""
    from oracle import reach_oracle as orc          # bench.py may use the oracle to prepare inputs / CPU arm only
    from reachability_jl_b200 import workloads as wl
    w = {"C1": wl.building, "C3": wl.fom}[name]()
    X0 = orc.Box(w.c, w.r)
    U = orc.Box(w.cU, w.rU)
    Phi = orc.exp_Adelta(w.A, w.delta)
    P2 = orc.phi12_series(np.abs(w.A), w.delta)[2]
    D = orc.discretize_box(w.A, X0, w.delta, w.B, U, Phi=Phi, Phi2abs=P2)
    N = orc.n_steps_bffpsv18(w.T, w.delta)
    return w, D, N


def box_workload_label(w, N):
    """`config.workload` of a BFFPSV18 line -- the SAME string on the GPU arm and on the reference arm (`same_config`)."""
    return (f"{w.name}: n={w.n}, m={w.m}, BFFPSV18 {len(w.partition)} blocks of dim {len(w.partition[0])}, delta={w.delta}, "
            f"T={w.T}, N={N} reach sets")


def zono_workload_label(w, N):
    return f"{w.name}: n={w.n}, m={w.m}, GLGM06 max_order={w.max_order}, delta={w.delta}, T={w.T}, N={N} reach sets"


def shard_rows(n, partition, rank, world):
    """Contiguous row blocks aligned to partition blocks, balanced by row count."""
    from reachability_jl_b200 import sharding
    lo, hi = sharding.shard_blocks([len(b) for b in partition], rank, world)
    return np.array([l for b in partition[lo:hi] for l in b], dtype=np.int32)


def box_algorithmic_flops_per_step(n, R, m):
    """2 R n^2 (power) + 4 R n (c, r) + 2 R n m + 4 R m + 2 R n (inputs)."""
    return 2.0 * R * n * n + 4.0 * R * n + 2.0 * R * n * m + 4.0 * R * m + 2.0 * R * n


class _Disc:
    """Attribute view of the dict `Context.discretize_box` returns (same field names as the oracle's result)."""

    def __init__(self, d):
        self.__dict__.update(d)


def run_bffpsv18(args, rank, world, dist):
    import torch
    import reachability_jl_b200 as rb
    import math
    from reachability_jl_b200 import workloads as wl
    big = args.workload == "C4"
    # The inputs of the loop (Phi, box of Omega0, discretised input) come from the library's own device discretisation,
    # outside every timed region: nothing under oracle/ is on this arm's path (the CPU legs below use it).
    w = {"C1": wl.building, "C3": wl.fom, "C4": wl.mna5}[args.workload]()
    D = None
    N = int(w.extra["N"]) if big else int(math.ceil(w.T / w.delta))          # reach.jl:54
    if args.flow_steps:
        N = args.flow_steps
    n, m = w.n, w.m
    rows = shard_rows(n, w.partition, rank, world)
    R = len(rows)
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    stream = torch.cuda.Stream(device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{dev}")    # > 126 MB L2
    with torch.cuda.stream(stream):
        ctx = rb.Context(dev, stream=stream.cuda_stream)
        t0 = time.perf_counter()
        D = _Disc(ctx.discretize_box(w.A, w.delta, w.c, w.r, w.B, w.cU, w.rU))
        disc_s = time.perf_counter() - t0
        # ---- device-resident arm: inputs uploaded once, K complete flowpipes timed ------------------------
        plan = ctx.boxplan(D.Phi, D.c0, D.r0, N, D.MU, D.cU, D.rU, D.rpsi, rows)
        # C4: a full flowpipe is minutes of DMMA work, so the warm-up passes run a truncated flowpipe of the same problem
        warm_N = min(N, 48) if big else N
        warm_plan = ctx.boxplan(D.Phi, D.c0, D.r0, warm_N, D.MU, D.cU, D.rU, D.rpsi, rows) if warm_N != N else plan
        gathered = mine = None
        if world > 1:
            sizes = [len(shard_rows(n, w.partition, r, world)) for r in range(world)]
            width = max(sizes) * N              # shards padded to the largest one: a plain equal-size NCCL all-gather
            mine = torch.zeros((2, width), dtype=torch.float64, device=f"cuda:{dev}")
            gathered = [torch.empty((2, width), dtype=torch.float64, device=f"cuda:{dev}") for _ in sizes]

        def one_flowpipe():
            flush.zero_()                      # L2 flush between timed iterations
            plan.run()
            if world > 1:
                cptr, rptr = plan.device_results()
                mine[0, :R * N].copy_(_as_tensor(cptr, R * N, dev))
                mine[1, :R * N].copy_(_as_tensor(rptr, R * N, dev))
                dist.all_gather(gathered, mine)   # "gather only at the end" (NCCL over NVLink)

        for _ in range(args.warmup):
            if warm_plan is plan:
                one_flowpipe()
            else:
                flush.zero_()
                warm_plan.run()
        if world > 1 and warm_plan is not plan:      # NCCL warm-up of the end-of-run gather
            dist.all_gather(gathered, mine)
        sampler = ClockSampler(dev)
        launches0 = ctx.launch_count()
        _barrier(dist, world)
        torch.cuda.synchronize()
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        gemm_ms = gemm_flops = 0.0
        gemm_launches = 0
        for _ in range(args.steps):
            one_flowpipe()
            t = plan.timing()
            gemm_ms += t["gemm_ms"]
            gemm_flops += t["gemm_flops"]
            gemm_launches += t["gemm_launches"]
        ev1.record(stream)
        torch.cuda.synchronize()
        _barrier(dist, world)
        clocks = sampler.stop()
        ms = ev0.elapsed_time(ev1)
        launches = ctx.launch_count() - launches0 + args.steps   # + the L2-flush memsets
        ms = _max_over_ranks(dist, world, ms, dev)
        # ---- end-to-end arm: host buffers in, host buffers out, through the one-shot C-ABI call -------------
        hPhi = pinned((n, n)); hPhi[:] = D.Phi
        hMU = pinned((n, m)); hMU[:] = D.MU
        hc = pinned((R, N)); hr = pinned((R, N))
        vecs = {k: pinned((len(v),)) for k, v in dict(c0=D.c0, r0=D.r0, cU=D.cU, rU=D.rU, rpsi=D.rpsi).items()}
        for k, v in dict(c0=D.c0, r0=D.r0, cU=D.cU, rU=D.rU, rpsi=D.rpsi).items():
            vecs[k][:] = v
        e2e_steps = 0 if args.no_e2e else (1 if big else max(1, min(args.steps, 3)))
        for _ in range(1 if (e2e_steps and not big) else 0):
            ctx.bffpsv18_boxes_into(hPhi, vecs["c0"], vecs["r0"], N, hMU, vecs["cU"], vecs["rU"], vecs["rpsi"], rows, hc, hr)
        _barrier(dist, world)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.bffpsv18_boxes_into(hPhi, vecs["c0"], vecs["r0"], N, hMU, vecs["cU"], vecs["rU"], vecs["rpsi"], rows, hc, hr)
        torch.cuda.synchronize()
        e2e_s = _max_over_ranks(dist, world, time.perf_counter() - t0, dev)
        h2d = 8 * (n * n + n * m + 3 * n + 2 * m) + 4 * R
        d2h = 16 * R * N
        # solve()-shaped call: device discretisation (A, X0, B, U from the host) + the flowpipe, host results out
        solve_s = None
        if e2e_steps and not big:
            hA = pinned((n, n)); hA[:] = w.A

            def one_solve():
                Dd = ctx.discretize_box(hA, w.delta, w.c, w.r, w.B, w.cU, w.rU)
                ctx.bffpsv18_boxes_into(Dd["Phi"], Dd["c0"], Dd["r0"], N, Dd["MU"], Dd["cU"], Dd["rU"], Dd["rpsi"], rows, hc, hr)

            one_solve()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            one_solve()
            torch.cuda.synchronize()
            solve_s = _max_over_ranks(dist, world, time.perf_counter() - t0, dev)
        if warm_plan is not plan:
            warm_plan.close()
        plan.close()
        ctx.close()
    lazy_obj = None
    rider = getattr(args, "rider", False)
    if big and rank == 0 and world == 1 and not rider:
        lazy_obj = run_c4_lazy(w, dev, N)
    if big and rank == 0 and world == 1 and not args.no_cpu_baseline and not rider:
        args._c4_cpu = cpu_baseline_c4(w, D, N)
    if rank != 0:
        return None
    flow_steps = N - 1
    value = args.steps * flow_steps / (ms * 1e-3)
    alg_flops = box_algorithmic_flops_per_step(n, R, m) * flow_steps * args.steps
    achieved = alg_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    out = {
        "metric": METRIC, "value": value, "unit": "flowpipe steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": box_workload_label(w, N), "bench_step": "one complete flowpipe (all N reach sets)",
                   "rows_per_gpu": R, "parallelism": f"row-block sharding x{world}, Phi replicated" if world > 1 else "single GPU",
                   "l2": "256 MB L2 flush between timed iterations; " + ("Phi (953 MB) and the power stacks exceed L2" if big else
                                                                         "power stacks 2 x 132 MB exceed L2"),
                   "device_discretize_s": disc_s,
                   **({"warmup_note": f"warm-up passes run the first {warm_N} reach sets of the same problem (a full flowpipe is "
                                      f"{2.0 * R * n * n * (N - 1) / 1e12:.0f} TFLOP on this rank)"} if big else {})},
        "e2e": {"value": e2e_steps * flow_steps / e2e_s if e2e_steps else None, "unit": "flowpipe steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps, "api": "reach_bffpsv18_boxes (host buffers, pinned)",
                "with_discretize": {"value": flow_steps / solve_s if solve_s else None, "unit": "flowpipe steps/s",
                                    "api": "reach_discretize_box + reach_bffpsv18_boxes: N / t_solve of SURVEY 8(d), expm and "
                                           "Phi2(|A|) on the device included"} if not big else None},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                     "frac": achieved / FP64_DMMA_PEAK_TFLOPS if achieved else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one steady-state launch (21 stacked row blocks: M=21168,
                     # N=1008, K=1006) from the `ncu --set full` capture profiles/r02/ncu_c3_power_gemm_final.txt: 182.4 MB +
                     # 146.3 MB (algorithmic: 171 MB stack read + 8 MB Phi^s + 171 MB stack write, part of which stays in L2)
                     "traffic": 328.7e6 if (w.n == 1006 and world == 1) else None,
                     "kernel": "gemm_dmma_kernel (stacked power GEMM S·[Phi^s | dB | c0])",
                     "launches": int(gemm_launches), "kernel_ms_total": gemm_ms,
                     "kernel_share_of_step": gemm_ms / ms,
                     "peak_source": "FP64 DMMA register-resident loop measured on this pool's B200 "
                                    "(profiles/r01/fp64_peak_b200.jsonl); MEASURED_PEAKS.json holds no FP64 figure "
                                    "(tcgen05 has no f64 kind, so its bf16 peak does not apply)",
                     "executed_flops": gemm_flops},
    }
    if lazy_obj is not None:
        out["lazy_expm_rows"] = lazy_obj
    return out


def run_c4_lazy(w, dev, N, nrows=31):
    """C4 with `exp_method = "lazy"`: sparse A on the device, sparse discretisation, `nrows` rows of interest
    advanced by the action of e^{δAᵀ}.  Rides along on the C4 line; timed with CUDA events inside the library."""
    import scipy.sparse as sp
    import reachability_jl_b200 as rb
    n = w.n
    rows = np.unique(np.concatenate([np.arange(10), np.linspace(0, n - 1, nrows - 10).astype(int)]))[:nrows].astype(np.int32)
    with rb.Context(dev) as ctx:
        S = ctx.sparse(sp.csc_matrix(w.A))
        S.discretize_box(w.delta, w.c, w.r, w.B, w.cU, w.rU)
        t0 = time.perf_counter()
        Ds = S.discretize_box(w.delta, w.c, w.r, w.B, w.cU, w.rU)
        disc_s = time.perf_counter() - t0
        best = None
        for _ in range(3):
            S.bffpsv18_boxes(w.delta, Ds["c0"], Ds["r0"], N, Ds["MU"], Ds["cU"], Ds["rU"], Ds["rpsi"], rows)
            t = S.timing()
            best = t if best is None or t["run_ms"] < best["run_ms"] else best
        # all n rows with the lazy exponential (the dense DMMA recurrence of the main line computes the same boxes): the
        # block no longer fits in L2, every Taylor term streams it through HBM; a truncated flowpipe is enough to time it
        N_all = min(N, 120)
        S.bffpsv18_boxes(w.delta, Ds["c0"], Ds["r0"], 8, Ds["MU"], Ds["cU"], Ds["rU"], Ds["rpsi"], None)
        S.bffpsv18_boxes(w.delta, Ds["c0"], Ds["r0"], N_all, Ds["MU"], Ds["cU"], Ds["rU"], Ds["rpsi"], None)
        ta = S.timing()
        S.close()
    all_rows = {"value": (N_all - 1) / (ta["run_ms"] * 1e-3), "unit": "flowpipe steps/s", "rows_of_interest": int(n),
                "reach_sets_timed": N_all, "algorithmic_gbs": ta["bytes"] / (ta["run_ms"] * 1e-3) / 1e9,
                "note": "same boxes as the dense recurrence of this line's `value` (parity: tests/test_lazy_sparse_gpu.py), "
                        "2 x 22 x nnz x n flop per step instead of 2 n^3"}
    return {"value": (N - 1) / (best["run_ms"] * 1e-3), "unit": "flowpipe steps/s", "rows_of_interest": int(len(rows)),
            "all_rows": all_rows,
            "sparse_discretize_s": disc_s, "products_per_step": best["products"] / (N - 1),
            "algorithmic_gbs": best["bytes"] / (best["run_ms"] * 1e-3) / 1e9,
            "api": "reach_sparse_create + reach_discretize_box_sparse + reach_bffpsv18_boxes_lazy (best of 3, CUDA events)"}


def run_c4_lazy_sharded(rank, world, dist, N=160):
    """Rider of the default line: C4 (n = 10913, sparse A) with `exp_method = "lazy"` and ALL rows of interest, the rows sharded
    over the ranks (no exchange: every row of e^{k delta A} advances on its own).  Same boxes as the dense
    DMMA recurrence of the `c4_mna5` rider (tests/test_lazy_sparse_gpu.py), 2 x 22 x nnz x rows flop per step instead of 2 n^2 rows."""
    import scipy.sparse as sp
    import torch
    import reachability_jl_b200 as rb
    from reachability_jl_b200 import workloads as wl
    w = wl.mna5()
    n = w.n
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    rows = np.array(shard_rows(n, w.partition, rank, world), dtype=np.int32)
    with rb.Context(dev) as ctx:
        S = ctx.sparse(sp.csc_matrix(w.A))
        Ds = S.discretize_box(w.delta, w.c, w.r, w.B, w.cU, w.rU)
        S.bffpsv18_boxes(w.delta, Ds["c0"], Ds["r0"], 8, Ds["MU"], Ds["cU"], Ds["rU"], Ds["rpsi"], rows)      # warm-up
        _barrier(dist, world)
        S.bffpsv18_boxes(w.delta, Ds["c0"], Ds["r0"], N, Ds["MU"], Ds["cU"], Ds["rU"], Ds["rpsi"], rows)
        t = S.timing()
        S.close()
    ms = _max_over_ranks(dist, world, t["run_ms"], dev)
    if rank != 0:
        return None
    return {"value": (N - 1) / (ms * 1e-3), "unit": "flowpipe steps/s", "ms_per_step": ms, "rows_per_gpu": int(len(rows)),
            "reach_sets_timed": N, "scaling": "strong",
            "config": {"workload": f"C4-mna5: n={n}, m={w.m}, BFFPSV18 {n} blocks of dim 1, delta={w.delta}, exp_method=lazy (sparse A, "
                                   f"nnz={int(np.count_nonzero(w.A))}), first {N} reach sets, all rows",
                       "parallelism": "single GPU" if world == 1 else f"rows sharded x{world}, sparse A replicated, no collective"},
            "roofline": {"bound": "hbm", "achieved": t["bytes"] / (t["run_ms"] * 1e-3) / 1e9, "peak": _peaks()[0], "unit": "GB/s",
                         "frac": t["bytes"] / (t["run_ms"] * 1e-3) / 1e9 / _peaks()[0],
                         "kernel": "spmm_taylor_kernel (Horner terms of the action of e^{delta A^T} on 256-row panels; algorithmic "
                                   "bytes 12 nnz + 24 n R per term, rank 0's rows; the panels are sized to stay in L2)"},
            "api": "reach_sparse_create + reach_discretize_box_sparse + reach_bffpsv18_boxes_lazy (CUDA events inside the library, max over ranks)"}


def build_zono_workload():
    """Discretised GLGM06 problem of config C2 (ISS-shaped, n=270, max_order 10) on the HOST, for the CPU legs only: Omega0
    (already reduce_order'ed, GLGM06/post.jl:20), V and Phi."""
    from oracle import reach_oracle as orc
    from reachability_jl_b200 import workloads as wl
    w = wl.iss()
    Phi = orc.exp_Adelta(w.A, w.delta)
    P2 = orc.phi12_series(np.abs(w.A), w.delta)[2]
    D = orc.discretize_zonotope(w.A, orc.Box(w.c, w.r), w.delta, w.B, orc.Box(w.cU, w.rU), Phi=Phi, Phi2abs=P2)
    Om = orc.reduce_order(D.Omega0, w.max_order)
    N = orc.n_steps_glgm06(w.T, w.delta)
    return w, D, Om, N


def glgm06_algorithmic_bytes_per_step(n, p0, pV, r):
    """SURVEY.md section 8(d), GLGM06 reduce x2: per reduction 8 n (p_in [score] + p_in [gather/box] + p_out [write])
    at the steady state p(W) = r n."""
    pW = r * n
    return 8.0 * n * ((p0 + pW) * 2 + r * n) + 8.0 * n * ((pW + pV) * 2 + r * n)


def run_glgm06(args, rank, world, dist, e2e=True):
    """C2 on one GPU; with N > 1 every rank runs a replica (column sharding of one system exists, sharding.run_sharded,
    but is latency-bound at n = 270: SURVEY.md section 8e)."""
    import torch
    import reachability_jl_b200 as rb
    from reachability_jl_b200 import api, workloads as wl
    w = wl.iss()
    N = args.flow_steps or int(round(w.T / w.delta))                         # GLGM06/post.jl:12
    n, r = w.n, w.max_order
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    stream = torch.cuda.Stream(device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{dev}")
    with torch.cuda.stream(stream):
        ctx = rb.Context(dev, stream=stream.cuda_stream)
        # discretised problem from the product path itself (device expm / Phi2 + the host-side zonotope bookkeeping of
        # api.discretize_zonotope, device reduce_order of Omega0: GLGM06/post.jl:15-20), outside every timed region
        Phi, (c0z, G0z), Vz = api.discretize_zonotope(ctx, w.A, w.delta, w.c, w.r, w.B, (w.cU, w.rU))
        G0red, _ = ctx.reduce_order(c0z, G0z, r)
        D = _Disc(dict(Phi=Phi, V=_Disc(dict(c=Vz[0], G=Vz[1], ngens=Vz[1].shape[1]))))
        Om = _Disc(dict(c=c0z, G=G0red, ngens=G0red.shape[1]))
        p0, pV = Om.ngens, D.V.ngens
        fp = ctx.glgm06(D.Phi, Om.c, Om.G, N, r, D.V.c, D.V.G, 0)

        def one_flowpipe():
            flush.zero_()
            fp.run()

        for _ in range(args.warmup):
            one_flowpipe()
        sampler = ClockSampler(dev)
        launches0 = ctx.launch_count()
        _barrier(dist, world)
        torch.cuda.synchronize()
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        tt = dict(gemm_ms=0.0, gemm_flops=0.0, reduce_ms=0.0, gemm_launches=0, chain_ms=0.0)
        for _ in range(args.steps):
            one_flowpipe()
            t = fp.timing()
            for key in tt:
                tt[key] += t[key]
        ev1.record(stream)
        torch.cuda.synchronize()
        _barrier(dist, world)
        clocks = sampler.stop()
        ms = _max_over_ranks(dist, world, ev0.elapsed_time(ev1), dev)
        launches = ctx.launch_count() - launches0 + args.steps
        p = fp.ngens()
        pmax = int(p.max())
        map_info = fp.map_info()
        # diagnostics pass outside the timed region: the same flowpipe with every kernel alone on one stream
        fp.set_serial(True)
        fp.run()
        alone = fp.timing()
        fp.set_serial(False)
        fp.close()
        # Phi of this workload is structurally sparse (decoupled modes), so the library maps the generators with an ELL
        # sparse product.  The dense DMMA GEMM path (what a dense Phi gets) is timed too, same flowpipe, same results:
        dense = None
        if map_info["sparse_levels"] > 0:
            os.environ["REACH_GLGM06_DENSE"] = "1"
            try:
                fd = ctx.glgm06(D.Phi, Om.c, Om.G, N, r, D.V.c, D.V.G, 0)
            finally:
                del os.environ["REACH_GLGM06_DENSE"]
            for _ in range(2):
                flush.zero_()
                fd.run()
            dtt = dict(gemm_ms=0.0, gemm_flops=0.0, reduce_ms=0.0, chain_ms=0.0)
            evd0, evd1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            evd0.record(stream)
            for _ in range(args.steps):
                flush.zero_()
                fd.run()
                t = fd.timing()
                for key in dtt:
                    dtt[key] += t[key]
            evd1.record(stream)
            torch.cuda.synchronize()
            dms = evd0.elapsed_time(evd1)
            fd.set_serial(True)
            fd.run()
            dalone = fd.timing()
            fd.close()
            dense = {"value": world * args.steps * (N - 1) / (dms * 1e-3), "unit": "flowpipe steps/s", "ms_per_step": dms / args.steps,
                     "note": "REACH_GLGM06_DENSE=1: linear_map(Phi^w, .) as the left-multiplication DMMA GEMM (the path of a dense Phi)",
                     "roofline": {"bound": "tensor", "achieved": dtt["gemm_flops"] / (dtt["gemm_ms"] * 1e-3) / 1e12,
                                  "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                                  "frac": dtt["gemm_flops"] / (dtt["gemm_ms"] * 1e-3) / 1e12 / FP64_DMMA_PEAK_TFLOPS,
                                  "traffic": 342.2e6, "kernel": "gemm_dmma_kernel<3,true> in situ (64-step launches on the SMs the "
                                  "partition leaves to it)", "kernel_ms_total": dtt["gemm_ms"], "kernel_share_of_step": dtt["gemm_ms"] / dms,
                                  "alone": {"gemm_tflops": dalone["gemm_flops"] / (dalone["gemm_ms"] * 1e-3) / 1e12,
                                            "gemm_ms": dalone["gemm_ms"], "numerics_ms": dalone["numerics_ms"],
                                            "chain_ms": dalone["chain_ms"], "run_ms": dalone["run_ms"]}}}
        # ---- end to end: host buffers in, all N zonotopes back in host memory (what reach_inhomog! fills) ----
        e2e_obj = None
        if e2e and not args.no_e2e:
            hC = pinned((n, N))
            hG = pinned((n, pmax, N))
            hPhi = pinned((n, n)); hPhi[:] = D.Phi
            hG0 = pinned((n, p0)); hG0[:] = Om.G
            hGV = pinned((n, pV)); hGV[:] = D.V.G

            def one_e2e():
                f2 = ctx.glgm06(hPhi, Om.c, hG0, N, r, D.V.c, hGV, 0)
                f2.run_fetch(hC, hG)
                f2.close()

            one_e2e()
            e2e_steps = max(1, min(args.steps, 2))
            _barrier(dist, world)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                one_e2e()
            torch.cuda.synchronize()
            e2e_s = _max_over_ranks(dist, world, time.perf_counter() - t0, dev)
            # solve() + project(sol, plot_vars): the flowpipe stays in HBM, every reach set is projected on the device
            # (reach_flowpipe_project) and only the 2 x N boxes cross PCIe
            P2 = np.zeros((2, n)); P2[0, 0] = 1.0; P2[1, n // 2] = 1.0

            def one_proj():
                f2 = ctx.glgm06(hPhi, Om.c, hG0, N, r, D.V.c, hGV, 0)
                f2.run()
                out = f2.project(P2)
                f2.close()
                return out

            one_proj()
            one_proj()
            proj_steps = 15                     # a solve + project call is ~11 ms: more of them than of the 1.3 s full fetch
            torch.cuda.synchronize()
            proj_call_ms = []
            t0 = time.perf_counter()
            for _ in range(proj_steps):
                tc = time.perf_counter()
                one_proj()
                proj_call_ms.append(round((time.perf_counter() - tc) * 1e3, 2))
            torch.cuda.synchronize()
            proj_s = _max_over_ranks(dist, world, time.perf_counter() - t0, dev)
            e2e_obj = {"value": world * e2e_steps * (N - 1) / e2e_s, "unit": "flowpipe steps/s",
                       "h2d_bytes_per_step": 8 * (n * n + n * (p0 + pV) + 2 * n),
                       "d2h_bytes_per_step": int(8 * n * (int(p.sum()) + N)), "steps": e2e_steps,
                       "api": "reach_glgm06_create + reach_glgm06_run_fetch (pinned host buffers; every zonotope of the "
                              "flowpipe copied back, batch by batch while later batches compute, as reach_inhomog! materialises them)",
                       # one call in ten takes 20-100 ms instead of 11 on the bench host (host-side hiccups, the device work is
                       # the same): `value` is N-1 over the MEDIAN call, the total over all calls is kept beside it
                       "projected": {"value": world * (N - 1) / (float(np.median(proj_call_ms)) * 1e-3), "unit": "flowpipe steps/s",
                                     "steps": proj_steps, "value_from_total_time": world * proj_steps * (N - 1) / proj_s,
                                     "d2h_bytes_per_step": 2 * 2 * 8 * N, "call_ms": proj_call_ms,
                                     "api": "reach_glgm06_create + reach_glgm06_run + reach_flowpipe_project (solve + project(sol, vars): "
                                            "2-D boxes of every reach set computed on the device, only they are copied back)"}}
        ctx.close()
    if rank != 0:
        return None
    flow_steps = N - 1
    value = world * args.steps * flow_steps / (ms * 1e-3)
    hbm, hbm_src = _peaks()
    alg_bytes = glgm06_algorithmic_bytes_per_step(n, p0, pV, r) * flow_steps * args.steps
    red_s = tt["reduce_ms"] * 1e-3
    gemm_tf = tt["gemm_flops"] / (tt["gemm_ms"] * 1e-3) / 1e12 if tt["gemm_ms"] else None
    nbatches = int(tt["gemm_launches"])
    # REAL traffic of the HBM-side kernels (generator map + scores + the batched numerics), dram__bytes_read.sum +
    # dram__bytes_write.sum per 64-step batch from the `ncu --set full` captures (C2_REAL_BYTES_PER_STEP below), against the time
    # those kernels take ALONE (the serialised diagnostics pass): in the pipelined run they overlap each other and the chain.
    real_bytes = C2_REAL_BYTES_PER_STEP * flow_steps
    hbm_side_ms = alone["gemm_ms"] + alone["numerics_ms"] if map_info["sparse_levels"] > 0 else alone["numerics_ms"]
    real_gbs = real_bytes / (hbm_side_ms * 1e-3) / 1e9 if hbm_side_ms > 0 else None
    numerics = {"kernel": "HBM-side kernels of a flowpipe: generator map fused with the Girard scores (ell_left_score, TMA-staged) + batched "
                          "numerics of the two reductions (wbox_partial, wseq, rank_r, apply_r, finalize_r)",
                "bound": "hbm", "achieved": real_gbs, "peak": hbm, "unit": "GB/s",
                "frac": real_gbs / hbm if real_gbs else None, "peak_source": hbm_src,
                "traffic": C2_REAL_BYTES_PER_STEP * 64, "traffic_unit": "bytes per 64-step batch (ncu dram read + write, " + C2_REAL_BYTES_SOURCE + ")",
                "kernel_ms_alone": hbm_side_ms, "kernel_ms_total": tt["reduce_ms"],
                "algorithmic": {"bytes_per_flowpipe_step": glgm06_algorithmic_bytes_per_step(n, p0, pV, r),
                                "gbs_by_survey_8d_bytes": alg_bytes / red_s / 1e9 if red_s > 0 else None,
                                "note": "SURVEY 8(d) counts 40.9 MB per step (score pass + gather pass + write of both reductions); the "
                                        "archive/reference design moves 7.1 MB (W and R_k are never rewritten -- reach sets are lists of "
                                        "generator references into the archive, scores travel with the references, box generators are "
                                        "tags) -- an algorithmic saving, not bandwidth"},
                "note": "in the pipelined run the HBM-side kernels share the GPU with the latency-bound ones (segment sort, selection "
                        "chain, box recurrence); the dense n x p matrix of a reach set is materialised by the fetch / projection kernels"}
    chain = {"kernel": "selection chain: chain_decide + chain_fast (parallel, one warp per generator) + chain_fast_fill in the steady state, "
                       "chain_kernel (one persistent CTA, sequential merge) in the growth phase", "kernel_ms_total": tt["chain_ms"],
             "us_per_flowpipe_step": 1e3 * tt["chain_ms"] / (flow_steps * args.steps)}
    alone_obj = {"note": "one extra flowpipe with all kernels serialised on one stream (not timed into `value`)",
                 "map_ms": alone["gemm_ms"], "numerics_ms": alone["numerics_ms"], "chain_ms": alone["chain_ms"],
                 "numerics_gbs": glgm06_algorithmic_bytes_per_step(n, p0, pV, r) * flow_steps / (alone["numerics_ms"] * 1e-3) / 1e9,
                 "run_ms": alone["run_ms"]}
    map_bytes = 16.0 * n * (p0 + pV + 2) * flow_steps * args.steps          # read X, write Y: 8 n q each way per step
    if map_info["sparse_levels"] > 0:
        # structurally sparse Phi: no GEMM on the path; the HBM-bound numerics are the dominant kernels by bytes, the selection
        # chain is the critical path by latency
        map_obj = {"kind": "ell_sparse", "nnz_per_row": map_info["nnz_per_row"], "launches": nbatches,
                   "kernel": "ell_left_score_kernel (Phi^w in ELL form, thread per state row, 16 generators per CTA staged with one bulk copy in "
                             "and one out, Girard scores and row-support masks fused)",
                   "kernel_ms_total": tt["gemm_ms"], "algorithmic_gbs": map_bytes / (tt["gemm_ms"] * 1e-3) / 1e9 if tt["gemm_ms"] else None,
                   "note": "Phi of this workload (decoupled modes) and its powers have 2 non-zeros per row; detected at creation, "
                           "results equal the dense GEMM's (tests/test_glgm06_gpu.py::test_sparse_phi_map_equals_dense_gemm); "
                           "`dense_gemm_path` times the DMMA path on the same flowpipe"}
        roof = dict(numerics, alone=alone_obj, selection_chain=chain)
    else:
        map_obj = {"kind": "dense_dmma_gemm", "launches": nbatches}
        roof = {"bound": "tensor", "achieved": gemm_tf, "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                "frac": gemm_tf / FP64_DMMA_PEAK_TFLOPS if gemm_tf else None, "traffic": 342.2e6,
                "kernel": "gemm_dmma_kernel<3,true> (Phi^s * archive slots [G0|GV|c0|cV], left multiplication, tensor-map TMA), in situ",
                "launches": nbatches, "kernel_ms_total": tt["gemm_ms"], "kernel_share_of_step": tt["gemm_ms"] / ms,
                "alone": alone_obj, "algorithmic_flops_per_flowpipe_step": 2.0 * n * n * (p0 + pV + 2),
                "peak_source": "FP64 DMMA register-resident loop measured on this pool's B200 (profiles/r01/fp64_peak_b200.jsonl)",
                "second_kernel": numerics, "selection_chain": chain}
    return {
        "metric": METRIC, "value": value, "unit": "flowpipe steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": zono_workload_label(w, N), "bench_step": "one complete flowpipe (all N reach sets)",
                   "generators": f"p0={p0}, pV={pV}, {pmax} generators per reach set",
                   "parallelism": "single GPU" if world == 1 else f"{world} independent replicas (the column-sharded variant, sharding.run_sharded, "
                                                                     "exists but does not beat one GPU at n=270, DESIGN.md section 5)",
                   "l2": "256 MB L2 flush between timed iterations; the flowpipe written per iteration (11.7 GB) exceeds L2"},
        "e2e": e2e_obj,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roof,
        "generator_map": map_obj,
        **({"dense_gemm_path": dense} if dense is not None else {}),
    }


def cpu_sample_glgm06(D, Om, r, n_steps):
    from oracle import reach_oracle as orc
    t0 = time.perf_counter()
    orc.glgm06_reach_inhomog(Om, D.V, D.Phi, n_steps, r)
    return time.perf_counter() - t0


def cpu_baseline_glgm06(target_s=12.0):
    w, D, Om, N = build_zono_workload()
    per = cpu_sample_glgm06(D, Om, w.max_order, 31) / 30
    n_cpu = int(max(31, min(N, target_s / max(per, 1e-9))))
    t = cpu_sample_glgm06(D, Om, w.max_order, n_cpu)
    return {"value": (n_cpu - 1) / t, "unit": "flowpipe steps/s", "cores": cpu_threads(), "kind": "port",
            "sample": f"first {n_cpu} of {N} reach sets of {w.name} (loop only, Omega0/V/Phi supplied), numpy+BLAS oracle port of "
                      "reach_inhomog! + reduce_order", "host_cpus": os.cpu_count()}


def cpu_baseline_ensemble(w, cs, rs, N, members=24):
    """C5 on the host: the oracle port of reach_homog! (GLGM06/reach.jl:5-24), one member after the other like the reference's
    `solve` per split (single task; the 28 x 28 products are too small for BLAS threads to matter)."""
    from oracle import reach_oracle as orc
    Phi = orc.exp_Adelta(w.A, w.delta)
    P2 = orc.Phi2(np.abs(w.A), w.delta)
    zs = [orc.discretize_zonotope(w.A, orc.Box(cs[:, b], rs[:, b]), w.delta, rzg=False, Phi=Phi, Phi2abs=P2).Omega0 for b in range(members)]
    t0 = time.perf_counter()
    for z in zs:
        orc.glgm06_reach_homog(z, Phi, N, rzg=False)
    t = time.perf_counter() - t0
    return {"value": members * (N - 1) / t, "unit": "ensemble member flowpipe steps/s", "cores": 1, "kind": "port",
            "sample": f"{members} of the 4096 members, all {N} reach sets each (loop only, Omega0 and Phi supplied), numpy oracle port of "
                      "reach_homog!, one member after the other like solve() per split", "host_cpus": os.cpu_count()}


def run_ensemble(args, rank, world, dist):
    """C5: batched homogeneous GLGM06 ensemble of a heli28-shaped system (n = 28).  512 members per GPU (the 4096
    splits of BASELINE.json's config on 8 GPUs), members sharded by rank, no collective: weak scaling."""
    import torch
    import reachability_jl_b200 as rb
    from reachability_jl_b200 import api, workloads as wl, sharding
    w = wl.heli28()
    n = w.n
    per_gpu = 512
    total = per_gpu * world
    N = args.flow_steps or int(round(w.T / w.delta))                      # GLGM06/post.jl:12
    cs, rs = wl.split_box(w.c, w.r, [4, 4, 4, 4, 2, 2, 2, 2])           # 4096 sub-boxes of X0
    sl = sharding.shard_batch(total, rank, world)
    cs, rs = cs[:, sl], rs[:, sl]
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    stream = torch.cuda.Stream(device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{dev}")
    with torch.cuda.stream(stream):
        ctx = rb.Context(dev, stream=stream.cuda_stream)
        # every member's Omega0 from the product path (device discretisation + host-side zonotope bookkeeping), untimed
        c0s = G0s = Phi = None
        for b in range(per_gpu):
            Phi, (c0b, G0b), _ = api.discretize_zonotope(ctx, w.A, w.delta, cs[:, b], rs[:, b], remove_zero_generators=False)
            if c0s is None:
                p = G0b.shape[1]
                c0s = np.zeros((n, per_gpu), order="F")
                G0s = np.zeros((n, p, per_gpu), order="F")
            c0s[:, b], G0s[:, :, b] = c0b, G0b
        ens = ctx.ensemble(Phi, c0s, G0s, N)

        def one():
            flush.zero_()
            ens.run()

        for _ in range(args.warmup):
            one()
        sampler = ClockSampler(dev)
        launches0 = ctx.launch_count()
        _barrier(dist, world)
        torch.cuda.synchronize()
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        kern_ms = 0.0
        for _ in range(args.steps):
            one()
            kern_ms += ens.timing()["run_ms"]
        ev1.record(stream)
        torch.cuda.synchronize()
        _barrier(dist, world)
        clocks = sampler.stop()
        ms = _max_over_ranks(dist, world, ev0.elapsed_time(ev1), dev)
        launches = ctx.launch_count() - launches0 + args.steps
        lo, hi = ens.boxes(N)
        assert np.all(np.isfinite(lo)) and np.all(hi >= lo)
        ens.close()
        # ---- end to end: the members' Omega0 from pinned host memory, the flowpipe, and the box hull of every member at every
        #      step back in pinned host memory (the zonotopes themselves -- 19.7 GB per GPU -- stay in HBM)
        e2e_obj = None
        if not args.no_e2e:
            hc0 = pinned((n, per_gpu)); hc0[:] = c0s
            hG0 = pinned((n, p, per_gpu)); hG0[:] = G0s
            hPhi = pinned((n, n)); hPhi[:] = Phi
            hlo = pinned((n, per_gpu, N)); hhi = pinned((n, per_gpu, N))

            def one_e2e():
                e2 = ctx.ensemble(hPhi, hc0, hG0, N)
                e2.run()
                e2.boxes_all(hlo, hhi)
                e2.close()

            one_e2e()
            e2e_steps = max(1, min(args.steps, 3))
            _barrier(dist, world)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                one_e2e()
            torch.cuda.synchronize()
            e2e_s = _max_over_ranks(dist, world, time.perf_counter() - t0, dev)
            e2e_obj = {"value": e2e_steps * (N - 1) * total / e2e_s, "unit": "ensemble member flowpipe steps/s",
                       "h2d_bytes_per_step": 8 * (n * n + n * per_gpu * (p + 1)), "d2h_bytes_per_step": 16 * n * per_gpu * N,
                       "steps": e2e_steps, "api": "reach_ensemble_create + reach_ensemble_run + reach_ensemble_boxes_all (pinned host "
                       "buffers; box hull of every member at every step copied back)"}
        ctx.close()
    if rank != 0:
        return None
    hbm, hbm_src = _peaks()
    steps = N - 1
    bytes_per_step = 16.0 * n * (p + 1) * per_gpu                         # read + write of one ensemble step
    real_bytes_per_step = 8.0 * n * (p + 1) * per_gpu                     # what the one-launch kernel moves: every slot is written once
    flops_per_step = 2.0 * n * n * (p + 1) * per_gpu                      # 2 n^2 (p+1) batch
    achieved = real_bytes_per_step * steps * args.steps / (kern_ms * 1e-3) / 1e9
    return {
        "metric": METRIC, "value": args.steps * steps * total / (ms * 1e-3), "unit": "ensemble member flowpipe steps/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C5-heli28 ensemble: n={n}, {p} generators, {per_gpu} members per GPU ({total} in total), "
                               f"homogeneous GLGM06, delta={w.delta}, N={N} reach sets per bench step",
                   "parallelism": f"members sharded x{world}, Phi replicated, no collective",
                   "l2": "256 MB L2 flush between timed iterations; the ensemble flowpipe (19.7 GB per GPU) exceeds L2"},
        "e2e": e2e_obj, "gpu_launches": int(launches), "clocks": clocks,
        # AI = 2 n^2 (p+1) flop / 8 n (p+1) B written = n / 4 = 7 flop/B against a machine balance of 37.06e12 / 6.55e12 = 5.7:
        # the one-launch kernel is bound by the FP64 tensor pipe (DMMA m8n8k4: 28 rows occupy 32), not by HBM -- both are reported
        "roofline": {"bound": "tensor", "achieved": flops_per_step * steps * args.steps / (kern_ms * 1e-3) / 1e12,
                     "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                     "frac": flops_per_step * steps * args.steps / (kern_ms * 1e-3) / 1e12 / FP64_DMMA_PEAK_TFLOPS,
                     "traffic": real_bytes_per_step * steps,
                     "kernel": "homog_small_dmma_kernel<4,7,NT> (whole flowpipe in ONE launch: Phi as DMMA A fragments in shared memory, "
                               "one CTA per SM with 19 warps, every slot written once by a bulk copy shared -> global and never re-read)",
                     "launches": args.steps, "kernel_ms_total": kern_ms, "kernel_share_of_step": kern_ms / ms,
                     "algorithmic_flops_per_ensemble_step": flops_per_step,
                     "executed_frac_with_m8_padding": flops_per_step * (32.0 / 28.0 if n == 28 else 1.0) * steps * args.steps / (kern_ms * 1e-3) / 1e12 / FP64_DMMA_PEAK_TFLOPS,
                     "peak_source": "FP64 DMMA register-resident loop measured on this pool's B200 (profiles/r01/fp64_peak_b200.jsonl)",
                     "hbm": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "peak_source": hbm_src,
                             "note": "REAL traffic = the writes (8 n (p+1) B per member and step; nothing is re-read); SURVEY 8(d)'s "
                                     "algorithmic read + write figure is twice that",
                             "algorithmic_bytes_per_ensemble_step": bytes_per_step}},
        **({} if getattr(args, "rider", False) or args.no_cpu_baseline or world > 1 else {"cpu_baseline": cpu_baseline_ensemble(w, wl.split_box(w.c, w.r, [4, 4, 4, 4, 2, 2, 2, 2])[0], wl.split_box(w.c, w.r, [4, 4, 4, 4, 2, 2, 2, 2])[1], N)}),
    }


def c2_flat_scalars(c2):
    """The C2 numbers as flat scalars for the end of the JSON line."""
    roof = c2.get("roofline") or {}
    chain = roof.get("selection_chain") or {}
    dense = c2.get("dense_gemm_path") or {}
    out = {"c2_value": c2.get("value"), "c2_unit": "flowpipe steps/s (n=270 GLGM06, N=2000)",
           "c2_e2e": (c2.get("e2e") or {}).get("value"), "c2_e2e_projected": ((c2.get("e2e") or {}).get("projected") or {}).get("value"),
           "c2_chain_us_per_step": chain.get("us_per_flowpipe_step"), "c2_real_hbm_frac": roof.get("frac"),
           "c2_dense_value": dense.get("value"), "c2_dense_frac": (dense.get("roofline") or {}).get("frac"),
           "c2_cpu_value": (c2.get("cpu_baseline") or {}).get("value")}
    return {k: v for k, v in out.items() if v is not None}


def run_glgm06_column_sharded(rank, world, dist, reps=3, dense=False):
    """Rider under --gpus N: ONE C2 system with its generator columns sharded over the N ranks (reach_glgm06_create_sharded,
    sharding.run_sharded): every rank maps, scores and gathers only its columns, the Girard selection runs replicated on the
    all-reduced scores, three small NCCL all-reduces per batch of steps."""
    import torch
    import reachability_jl_b200 as rb
    from reachability_jl_b200 import api, sharding, workloads as wl
    w = wl.iss()
    N = int(round(w.T / w.delta))
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        ctx = rb.Context(dev, stream=stream.cuda_stream)
        Phi, (c0, G0), V = api.discretize_zonotope(ctx, w.A, w.delta, w.c, w.r, w.B, (w.cU, w.rU))
        G0red, _ = ctx.reduce_order(c0, G0, w.max_order)
        if dense:            # the path of a dense Phi: linear_map(Phi^w, .) as the DMMA GEMM, which shards by columns
            os.environ["REACH_GLGM06_DENSE"] = "1"
        try:
            fp = ctx.glgm06_sharded(Phi, c0, G0red, N, w.max_order, V[0], V[1], rank, world)
        finally:
            os.environ.pop("REACH_GLGM06_DENSE", None)
        # phases in order on one stream: at n = 270 the map of a batch is as short as its selection / box phases, and issuing it
        # on a second stream (sharding.run_sharded(gemm_stream=...), what the n = 1024 rider does) measured slower on 2 GPUs
        # (sparse map 165 k against 190 k steps/s, dense map 125 k against 127 k)
        gemm_stream = None
        for _ in range(2):
            sharding.run_sharded(fp, dist, gemm_stream=gemm_stream)
        torch.cuda.synchronize()
        dist.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(reps):
            sharding.run_sharded(fp, dist, gemm_stream=gemm_stream)
        ev1.record(stream)
        torch.cuda.synchronize()
        ms = _max_over_ranks(dist, world, ev0.elapsed_time(ev1), dev)
        p = fp.ngens()
        fp.close()
        ctx.close()
    if rank != 0:
        return None
    return {"value": reps * (N - 1) / (ms * 1e-3), "unit": "flowpipe steps/s", "ms_per_flowpipe": ms / reps, "ranks": world,
            "generators_per_reach_set": int(p.max()), "scaling": "strong",
            "generator_map": "DMMA GEMM (dense Phi path, REACH_GLGM06_DENSE=1)" if dense else "ELL sparse map (Phi of C2 has 2 non-zeros per row)",
            "api": "reach_glgm06_create_sharded + reach_glgm06_shard_phase / _shard_buffer, NCCL sum all-reduces issued by "
                   "sharding.run_sharded on the library's stream (no host synchronisation inside the flowpipe)",
            "note": "at n = 270 the exchange is latency-bound (SURVEY 8e): this does not beat one GPU; it exists for large n * p"}


def run_glgm06_large(rank, world, dist, reps=3):
    """Rider: a large n * p GLGM06 system with a DENSE Phi (n = 1024, order 10, 4100 + 1027 mapped generators per step, 10240 per
    reach set; workloads.iss_mixed) -- the regime the generator-column sharding of SURVEY.md section 8e exists for.  Rank 0
    times the pipelined single-GPU engine (value, in-situ DMMA fraction of the generator map); under --gpus N all ranks then
    run the SAME system column-sharded with the software-pipelined phase machine (phase 0 of batch b + 1 on a second stream)."""
    import torch
    import reachability_jl_b200 as rb
    from reachability_jl_b200 import api, sharding, workloads as wl
    w = wl.iss_mixed()
    N = int(round(w.T / w.delta))
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    stream = torch.cuda.Stream(device=dev)
    out = None
    with torch.cuda.stream(stream):
        ctx = rb.Context(dev, stream=stream.cuda_stream)
        Phi, (c0, G0), V = api.discretize_zonotope(ctx, w.A, w.delta, w.c, w.r, w.B, (w.cU, w.rU))
        G0red, _ = ctx.reduce_order(c0, G0, w.max_order)
        if rank == 0:
            fp = ctx.glgm06(Phi, c0, G0red, N, w.max_order, V[0], V[1], 0)
            fp.run()
            best, t = 1e30, None
            for _ in range(reps):
                fp.run()
                ti = fp.timing()
                if ti["run_ms"] < best:
                    best, t = ti["run_ms"], ti
            pmax = int(fp.ngens().max())
            fp.close()
            tf = t["gemm_flops"] / (best * 1e-3) * 1e-12
            out = {"value": (N - 1) / (best * 1e-3), "unit": "flowpipe steps/s", "ms_per_flowpipe": best,
                   "config": {"workload": f"{w.name}: n={w.n}, m={w.m}, GLGM06 max_order={w.max_order}, delta={w.delta}, N={N} reach sets, "
                                          f"p0={G0red.shape[1]}, pV={V[1].shape[1]}, {pmax} generators per reach set, dense Phi"},
                   "roofline": {"bound": "tensor", "achieved": tf, "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                                "frac": tf / FP64_DMMA_PEAK_TFLOPS,
                                "kernel": "gemm_dmma_kernel<.,true> in situ: flops of the generator map / whole-flowpipe time "
                                          "(selection chain, scores and box sums overlapped on the SMs the GEMM leaves free)",
                                "gemm_ms_in_situ": t["gemm_ms"], "numerics_ms_in_situ": t["numerics_ms"], "chain_ms_in_situ": t["chain_ms"]}}
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            fs = ctx.glgm06_sharded(Phi, c0, G0red, N, w.max_order, V[0], V[1], rank, world)
            gs = torch.cuda.Stream(device=dev)
            for _ in range(2):
                sharding.run_sharded(fs, dist, gemm_stream=gs)
            torch.cuda.synchronize()
            dist.barrier()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            for _ in range(reps):
                sharding.run_sharded(fs, dist, gemm_stream=gs)
            ev1.record(stream)
            torch.cuda.synchronize()
            ms = _max_over_ranks(dist, world, ev0.elapsed_time(ev1), dev) / reps
            fs.close()
            if out is not None:
                out["column_sharded"] = {"value": (N - 1) / (ms * 1e-3), "unit": "flowpipe steps/s", "ms_per_flowpipe": ms, "ranks": world,
                                         "speedup_vs_one_gpu": out["ms_per_flowpipe"] / ms, "scaling": "strong",
                                         "api": "reach_glgm06_create_sharded + reach_glgm06_shard_phase_on / _shard_buffer; "
                                                "sharding.run_sharded(gemm_stream=...): three NCCL sum all-reduces per 64-step batch"}
        ctx.close()
    return out


def run_inprocess_multi(rank, world, dist):
    """Rider under --gpus N: ONE process (rank 0) drives all N GPUs through a multi-device context (reach_ctx_create_multi) and
    the one-shot host-buffer call reach_bffpsv18_boxes -- what a single Julia task gets through one `ccall`.  The other ranks
    wait at a barrier with their GPUs idle."""
    import torch
    import math
    import reachability_jl_b200 as rb
    from reachability_jl_b200 import workloads as wl
    out = None
    # the waiting ranks must leave their GPUs IDLE: an NCCL barrier is a kernel that spins on the device, and rank 0's work on
    # that device would be time-sliced against it (two processes on one GPU) -- so this rider synchronises on the host (gloo)
    os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")
    host_group = dist.new_group(backend="gloo")
    torch.cuda.synchronize()
    dist.barrier(group=host_group)
    if rank == 0:
        w = wl.fom()
        N = int(math.ceil(w.T / w.delta))
        n, m = w.n, w.m
        with rb.Context(list(range(world))) as ctx:
            D = ctx.discretize_box(w.A, w.delta, w.c, w.r, w.B, w.cU, w.rU)
            hPhi = pinned((n, n)); hPhi[:] = D["Phi"]
            hMU = pinned((n, m)); hMU[:] = D["MU"]
            hc = pinned((n, N)); hr = pinned((n, N))
            rows = np.arange(n, dtype=np.int32)
            call = lambda: ctx.bffpsv18_boxes_into(hPhi, D["c0"], D["r0"], N, hMU, D["cU"], D["rU"], D["rpsi"], rows, hc, hr)
            call()
            reps = 3
            t0 = time.perf_counter()
            for _ in range(reps):
                call()
            dt = time.perf_counter() - t0
        out = {"value": reps * (N - 1) / dt, "unit": "flowpipe steps/s", "devices": world, "steps": reps,
               "api": "reach_ctx_create_multi + reach_bffpsv18_boxes: one process, one handle, rows sharded over the devices inside "
                      "the library (host threads, no collective), pinned host buffers in and out",
               "h2d_bytes_per_step": world * 8 * (n * n + n * m + 3 * n + 2 * m), "d2h_bytes_per_step": 16 * n * N}
    dist.barrier(group=host_group)
    return out


def _as_tensor(ptr, count, dev):
    """Zero-copy torch view of a device buffer owned by the library."""
    import torch

    class _Arr:
        pass
    a = _Arr()
    a.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(a, device=f"cuda:{dev}")


def _barrier(dist, world):
    if world > 1:
        dist.barrier()


def _max_over_ranks(dist, world, x, dev):
    if world == 1:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{dev}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference loop on the host cores
# ---------------------------------------------------------------------------------------------------------

def cpu_sample_bffpsv18(D, n_steps):
    """Times `n_steps` reach sets of the dense reach_blocks! recurrence (oracle port) -> seconds."""
    from oracle import reach_oracle as orc
    t0 = time.perf_counter()
    orc.bffpsv18_reach(D, n_steps)
    return time.perf_counter() - t0


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [1])
    except Exception:
        return os.cpu_count() or 1


def cpu_baseline_c4(w, D, N, n_cpu=4):
    """C4: one reach set is a 2 n^3 = 2.6 TFLOP dgemm on the host, so the sample is the first few reach sets."""
    t = cpu_sample_bffpsv18(D, n_cpu)
    return {"value": (n_cpu - 1) / t, "unit": "flowpipe steps/s", "cores": cpu_threads(), "kind": "port",
            "sample": f"first {n_cpu} of {N} reach sets of {w.name} (loop only, Phi from the device discretisation), numpy+BLAS "
                      "oracle port of reach_blocks!", "host_cpus": os.cpu_count()}


def cpu_baseline_for(args, target_s=12.0):
    if args.workload == "C4":
        return getattr(args, "_c4_cpu", None)
    w, D, N = build_box_workload(args.workload)
    t_probe = cpu_sample_bffpsv18(D, 6)                 # 5 steps
    per = t_probe / 5
    n_cpu = int(max(6, min(N, target_s / max(per, 1e-9))))
    t = cpu_sample_bffpsv18(D, n_cpu)
    return {"value": (n_cpu - 1) / t, "unit": "flowpipe steps/s", "cores": cpu_threads(), "kind": "port",
            "sample": f"first {n_cpu} of {N} reach sets of {w.name} (loop only, Phi supplied), numpy+BLAS oracle port of "
                      "reach_blocks! without the lazy-set object overhead (=> faster than the Julia reference)",
            "host_cpus": os.cpu_count()}


def run_reference_glgm06(args, rank, world):
    if rank != 0:
        return None
    w, D, Om, N = build_zono_workload()
    per = cpu_sample_glgm06(D, Om, w.max_order, 31) / 30
    n_cpu = int(max(31, min(N, 4.0 / max(per, 1e-9))))
    for _ in range(args.warmup):
        cpu_sample_glgm06(D, Om, w.max_order, min(n_cpu, 20))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_sample_glgm06(D, Om, w.max_order, n_cpu)
    t = time.perf_counter() - t0
    value = args.steps * (n_cpu - 1) / t
    cores = cpu_threads()
    sample = f"each bench step = first {n_cpu} of {N} reach sets of {w.name} (loop only), oracle port, {cores} BLAS threads"
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "flowpipe steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": zono_workload_label(w, N), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "flowpipe steps/s", "cores": cores, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": "flowpipe steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def run_reference(args, rank, world):
    if rank != 0:
        return None
    if args.workload == "C2":
        return run_reference_glgm06(args, rank, world)
    if args.workload in ("C4", "C5"):
        return {"impl": "reference", "unavailable": f"{args.workload}: the CPU arm of this workload is the cpu_baseline object of the "
                "default arm (C4's inputs come from the device discretisation; a host expm of n=10913 alone takes minutes)"}
    w, D, N = build_box_workload(args.workload)
    per = cpu_sample_bffpsv18(D, 6) / 5
    n_cpu = int(max(6, min(N, 4.0 / max(per, 1e-9))))       # ~4 s per bench step
    for _ in range(args.warmup):
        cpu_sample_bffpsv18(D, min(n_cpu, 20))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_sample_bffpsv18(D, n_cpu)
    t = time.perf_counter() - t0
    value = args.steps * (n_cpu - 1) / t
    cores = cpu_threads()
    sample = f"each bench step = first {n_cpu} of {N} reach sets of {w.name} (loop only), oracle port, {cores} BLAS threads"
    return {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "flowpipe steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": box_workload_label(w, N), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "flowpipe steps/s", "cores": cores, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": "flowpipe steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="C3", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--no-secondary", action="store_true", help="skip the second headline workload (C2 GLGM06) summary")
    ap.add_argument("--no-riders", action="store_true", help="skip the C4 / C5 riders of the default line")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer arm")
    ap.add_argument("--flow-steps", type=int, default=0, help="profiling runs only: truncate the flowpipe to this many reach sets")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        out = run_reference(args, rank, world)
        if out is not None:
            print(json.dumps(out), flush=True)
        return
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    if args.workload == "C2":
        out = run_glgm06(args, rank, world, dist)
    elif args.workload == "C5":
        out = run_ensemble(args, rank, world, dist)
        args.no_cpu_baseline = True
    else:
        out = run_bffpsv18(args, rank, world, dist)
    flat = {}                       # flat scalars appended at the END of the line (they survive a truncated stdout tail)
    primary_c3 = args.workload == "C3" and not args.flow_steps
    if primary_c3 and not args.no_riders:
        # ---- riders of the default line: the other configurations of BASELINE.json, each reduced to a few scalars ----------
        # C4 (n = 10913, the configuration the 1 -> 8 GPU scaling target names): a truncated flowpipe of the same problem,
        # all rows sharded like the main line; C5: the ensemble, members sharded (weak scaling)
        r4 = run_bffpsv18(argparse.Namespace(**{**vars(args), "workload": "C4", "steps": 1, "warmup": 1, "flow_steps": 160,
                                                "no_e2e": True, "no_cpu_baseline": True, "rider": True}), rank, world, dist)
        r5 = run_ensemble(argparse.Namespace(**{**vars(args), "workload": "C5", "steps": 2, "warmup": 1, "flow_steps": 0,
                                                "no_e2e": world > 1, "no_cpu_baseline": True, "rider": True}), rank, world, dist)
        shard = run_glgm06_column_sharded(rank, world, dist) if world > 1 else None
        shard_dense = run_glgm06_column_sharded(rank, world, dist, dense=True) if world > 1 else None
        multi = run_inprocess_multi(rank, world, dist) if world > 1 else None
        large = run_glgm06_large(rank, world, dist)
        lazy4 = run_c4_lazy_sharded(rank, world, dist)
        if out is not None:
            out["riders"] = {"c4_mna5": {k: r4[k] for k in ("value", "unit", "ms_per_step", "config", "roofline", "scaling")},
                             "c5_ensemble": {k: r5[k] for k in ("value", "unit", "ms_per_step", "config", "roofline", "scaling", "e2e")}}
            flat.update(c4_value=r4["value"], c4_unit="flowpipe steps/s (n=10913, first 160 reach sets, all rows)",
                        c4_frac=r4["roofline"]["frac"], c5_value=r5["value"], c5_unit=r5["unit"],
                        c5_dmma_frac=r5["roofline"]["frac"], c5_real_hbm_frac=r5["roofline"]["hbm"]["frac"])
            if r5.get("e2e"):
                flat["c5_e2e"] = r5["e2e"]["value"]
            if shard is not None:
                out["riders"]["c2_column_sharded"] = shard
                flat["c2_column_sharded_value"] = shard["value"]
            if shard_dense is not None:
                shard_dense["note"] = ("dense-Phi path: the DMMA GEMM (the dominant cost on one GPU, see c2_dense_value of the N=1 "
                                       "line) shards by generator columns; compare with c2_dense_value, not with c2_value")
                out["riders"]["c2_column_sharded_dense"] = shard_dense
                flat["c2_column_sharded_dense_value"] = shard_dense["value"]
            if large is not None:
                out["riders"]["glgm06_dense_n1024"] = large
                flat.update(glgm06_n1024_dense_value=large["value"], glgm06_n1024_dense_frac=large["roofline"]["frac"])
                if "column_sharded" in large:
                    flat.update(glgm06_n1024_column_sharded_value=large["column_sharded"]["value"],
                                glgm06_n1024_column_sharded_speedup=large["column_sharded"]["speedup_vs_one_gpu"])
            if lazy4 is not None:
                out["riders"]["c4_lazy_expm_all_rows"] = lazy4
                flat.update(c4_lazy_value=lazy4["value"], c4_lazy_frac=lazy4["roofline"]["frac"])
            if multi is not None:
                out["riders"]["c3_one_process_multi_device"] = multi
                flat["c3_one_process_multi_device_e2e"] = multi["value"]
    if out is not None:
        if args.flow_steps:
            out["config"]["truncated"] = f"PROFILING RUN: flowpipe truncated to {args.flow_steps} reach sets"
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_glgm06() if args.workload == "C2" else cpu_baseline_for(args)
        if world == 1 and primary_c3 and not args.no_secondary:
            # the metric names two configurations: the n=270 GLGM06 one rides along as a summary object
            sec = run_glgm06(args, rank, world, dist)
            keep = ("value", "unit", "ms_per_step", "config", "e2e", "roofline", "gpu_launches", "generator_map", "dense_gemm_path")
            out["glgm06_c2"] = {k: sec[k] for k in keep if k in sec}
            if not args.no_cpu_baseline:
                out["glgm06_c2"]["cpu_baseline"] = cpu_baseline_glgm06()
            flat.update(c2_flat_scalars(out["glgm06_c2"]))
        if args.workload == "C2":
            flat.update(c2_flat_scalars(out))
        out.update(flat)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
