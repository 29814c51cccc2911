#!/bin/bash
# 1 GPU: full GPU suite, smoke, default bench line (cpu baseline + extra configs + full amplitude)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 700 python -m pytest tests -m gpu -q -rA --durations=8 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -5; grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu.log | head
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --dump-steps gpurun_out/r2_steps_default.json > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
echo "bench rc=$?" >> gpurun_out/r2_bench_default.err
cut -c1-300 gpurun_out/r2_bench_default.json; tail -3 gpurun_out/r2_bench_default.err
