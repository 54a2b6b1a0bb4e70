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

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

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
    """Discretised BFFPSV18 problem of config C1/C3/C4 (inputs of the hot loop).  Discretisation runs on the host
    here only to PREPARE the synthetic inputs (outside every timed region)."""
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


def shard_rows(n, partition, rank, world):
    """Contiguous row blocks aligned to partition blocks, balanced by row count."""
    nb = len(partition)
    lo = (nb * rank) // world
    hi = (nb * (rank + 1)) // world
    return np.array([l for b in partition[lo:hi] for l in b], dtype=np.int32)


def box_algorithmic_flops_per_step(n, R, m):
    """2 R n^2 (power) + 4 R n (c, r) + 2 R n m + 4 R m + 2 R n (inputs)."""
    return 2.0 * R * n * n + 4.0 * R * n + 2.0 * R * n * m + 4.0 * R * m + 2.0 * R * n


def run_bffpsv18(args, rank, world, dist):
    import torch
    import reachability_jl_b200 as rb
    w, D, N = build_box_workload(args.workload)
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
        # ---- device-resident arm: inputs uploaded once, K complete flowpipes timed ------------------------
        plan = ctx.boxplan(D.Phi, D.c0, D.r0, N, D.MU, D.cU, D.rU, D.rpsi, rows)
        gathered = None
        if world > 1:
            cptr, rptr = plan.device_results()
            mine = torch.stack([_as_tensor(cptr, R * N, dev), _as_tensor(rptr, R * N, dev)])
            sizes = [len(shard_rows(n, w.partition, r, world)) for r in range(world)]
            gathered = [torch.empty((2, sz * N), dtype=torch.float64, device=f"cuda:{dev}") for sz in sizes]

        def one_flowpipe():
            flush.zero_()                      # L2 flush between timed iterations
            plan.run()
            if world > 1:
                cptr, rptr = plan.device_results()
                mine = torch.stack([_as_tensor(cptr, R * N, dev), _as_tensor(rptr, R * N, dev)])
                dist.all_gather(gathered, mine)   # "gather only at the end" (NCCL over NVLink)

        for _ in range(args.warmup):
            one_flowpipe()
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
        e2e_steps = 0 if args.no_e2e else max(1, min(args.steps, 3))
        for _ in range(1 if e2e_steps else 0):
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
        plan.close()
        ctx.close()
    if rank != 0:
        return None
    flow_steps = N - 1
    value = args.steps * flow_steps / (ms * 1e-3)
    alg_flops = box_algorithmic_flops_per_step(n, R, m) * flow_steps * args.steps
    achieved = alg_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    out = {
        "metric": METRIC, "value": value, "unit": "flowpipe steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{w.name}: n={n}, m={m}, BFFPSV18 {len(w.partition)} blocks of dim "
                               f"{len(w.partition[0])}, delta={w.delta}, T={w.T}, N={N} reach sets per bench step",
                   "rows_per_gpu": R, "parallelism": f"row-block sharding x{world}, Phi replicated" if world > 1 else "single GPU",
                   "l2": "256 MB L2 flush between timed iterations; power stacks 2 x 132 MB exceed L2"},
        "e2e": {"value": e2e_steps * flow_steps / e2e_s if e2e_steps else None, "unit": "flowpipe steps/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps, "api": "reach_bffpsv18_boxes (host buffers, pinned)"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                     "frac": achieved / FP64_DMMA_PEAK_TFLOPS if achieved else None, "traffic": None,
                     "kernel": "gemm_dmma_kernel (stacked power GEMM S·[Phi^s | dB | c0])",
                     "launches": int(gemm_launches), "kernel_ms_total": gemm_ms,
                     "kernel_share_of_step": gemm_ms / ms,
                     "peak_source": "FP64 DMMA register-resident loop measured on this pool's B200 "
                                    "(profiles/r01/fp64_peak_b200.jsonl); MEASURED_PEAKS.json holds no FP64 figure "
                                    "(tcgen05 has no f64 kind, so its bf16 peak does not apply)",
                     "executed_flops": gemm_flops},
    }
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


def cpu_baseline_for(args, target_s=12.0):
    w, D, N = build_box_workload(args.workload)
    t_probe = cpu_sample_bffpsv18(D, 6)                 # 5 steps
    per = t_probe / 5
    n_cpu = int(max(6, min(N, target_s / max(per, 1e-9))))
    t = cpu_sample_bffpsv18(D, n_cpu)
    return {"value": (n_cpu - 1) / t, "unit": "flowpipe steps/s", "cores": cpu_threads(), "kind": "port",
            "sample": f"first {n_cpu} of {N} reach sets of {w.name} (loop only, Phi supplied), numpy+BLAS oracle port of "
                      "reach_blocks! without the lazy-set object overhead (=> faster than the Julia reference)",
            "host_cpus": os.cpu_count()}


def run_reference(args, rank, world):
    if rank != 0:
        return None
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
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{w.name}: n={w.n}, m={w.m}, BFFPSV18, delta={w.delta}, T={w.T}, N={N}", "sample": sample},
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
    ap.add_argument("--workload", default="C3", choices=["C1", "C3"])
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
    out = run_bffpsv18(args, rank, world, dist)
    if out is not None:
        if args.flow_steps:
            out["config"]["truncated"] = f"PROFILING RUN: flowpipe truncated to {args.flow_steps} reach sets"
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline_for(args)
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
