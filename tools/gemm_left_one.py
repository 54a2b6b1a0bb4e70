"""One shape of the left-multiplication GEMM (for ncu captures): n Q K reps."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import reachability_jl_b200 as rb
n, Q, K, reps = (int(a) for a in sys.argv[1:5])
rng = np.random.default_rng(0)
P = rng.standard_normal((n, K)); X = rng.standard_normal((K, Q))
with rb.Context(0) as ctx:
    Y, ms = ctx.dgemm_left(P, X, reps=reps)
    print("ms", ms, "TFLOP/s", 2.0 * n * Q * K / (ms * 1e-3) / 1e12)
