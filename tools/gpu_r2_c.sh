#!/bin/bash
# Round 2, call C: BFFPSV18 parity after the fused abs-sum epilogue, then C3 A/B (fused vs the separate streaming pass).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "bffpsv18 or api or multi_device or golden or c_client" > gpurun_out/pytest_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_c.log
tail -5 gpurun_out/pytest_c.log
timeout 600 python bench.py --no-secondary --no-cpu-baseline --steps 6 --warmup 3 > gpurun_out/bench_c3_fused.json 2> gpurun_out/bench_c3_fused.err; echo "rc=$?"
REACH_BFFPSV18_NO_FUSED_ABS=1 timeout 600 python bench.py --no-secondary --no-cpu-baseline --steps 6 --warmup 3 > gpurun_out/bench_c3_unfused.json 2> gpurun_out/bench_c3_unfused.err; echo "rc=$?"
python - <<'PY'
import json
for name in ("fused", "unfused"):
    try:
        d = json.loads(open(f"gpurun_out/bench_c3_{name}.json").read().strip().splitlines()[-1])
        r = d["roofline"]
        print(name, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(r["frac"], 4), "share", round(r["kernel_share_of_step"], 4), "ms/step", round(d["ms_per_step"], 2))
    except Exception as e:
        print(name, "failed", e)
PY
