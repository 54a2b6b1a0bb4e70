#!/bin/bash
# end-of-session validation on one box: GPU test suite, smoke(), default bench, reference arm
python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?" >> gpurun_out/final_bench.err
python bench.py --impl reference > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "ref rc=$?" >> gpurun_out/final_bench_reference.err
tail -3 gpurun_out/final_pytest.log; cat gpurun_out/final_smoke.log; grep -E "rc=" gpurun_out/final_bench.err; tail -c 400 gpurun_out/final_bench_reference.json
