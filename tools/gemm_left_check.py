"""GPU check of the left-multiplication (generator layout) DMMA GEMM against numpy (run under gpurun)."""
import json, sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import reachability_jl_b200 as rb
rng = np.random.default_rng(0)
shapes = [(4, 7, 4), (28, 87 * 64, 28), (64, 300, 64), (65, 300, 65), (130, 129, 130), (270, 1359, 270), (270, 1359 * 32, 270),
          (1006, 4000, 1006), (28, 4096 * 87, 28)]
with rb.Context(0) as ctx:
    for (n, Q, K) in shapes:
        P = rng.standard_normal((n, K))
        X = rng.standard_normal((K, Q))
        Y, ms = ctx.dgemm_left(P, X, reps=10)
        ref = P @ X
        err = float(np.max(np.abs(Y - ref)) / max(1.0, np.max(np.abs(ref))))
        bad = np.argwhere(np.abs(Y - ref) > 1e-9 * max(1.0, np.max(np.abs(ref))))
        print(json.dumps({"n": n, "Q": Q, "K": K, "rel_err": err, "ms": ms, "nbad": int(len(bad)),
                          "first_bad": bad[:4].tolist(), "tflops": 2.0 * n * Q * K / (ms * 1e-3) / 1e12 if ms > 0 else None}), flush=True)
