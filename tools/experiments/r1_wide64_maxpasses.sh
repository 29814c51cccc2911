#!/bin/bash
# Experiment round: stem kernel WIDE64 variant (TNB_STEM_WIDE64=1) and more stem passes (TNB_STEM_MAX_PASSES=32).
mkdir -p gpurun_out
timeout 200 python bench.py --dump-steps gpurun_out/steps_r2_default.json > gpurun_out/bench_r2_default.json 2> gpurun_out/bench_r2_default.err
echo "bench rc=$?" >> gpurun_out/bench_r2_default.err
TNB_STEM_WIDE64=1 timeout 200 python -m pytest tests -m gpu -q -rA -k "stem_tc or sycamore53 or regular3" > gpurun_out/pytest_wide64.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_wide64.log
TNB_STEM_WIDE64=1 timeout 150 python bench.py --no-cpu --dump-steps gpurun_out/steps_wide64.json > gpurun_out/bench_wide64.json 2> gpurun_out/bench_wide64.err
TNB_STEM_WIDE64=1 TNB_STEM_MAX_PASSES=32 timeout 150 python -m pytest tests -m gpu -q -rA -k "sycamore53" > gpurun_out/pytest_wide64_p32.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_wide64_p32.log
TNB_STEM_WIDE64=1 TNB_STEM_MAX_PASSES=32 timeout 150 python bench.py --no-cpu --dump-steps gpurun_out/steps_wide64_p32.json > gpurun_out/bench_wide64_p32.json 2> gpurun_out/bench_wide64_p32.err
TNB_STEM_MAX_PASSES=32 timeout 150 python bench.py --no-cpu --dump-steps gpurun_out/steps_p32.json > gpurun_out/bench_p32.json 2> gpurun_out/bench_p32.err
tail -3 gpurun_out/pytest_wide64.log; tail -3 gpurun_out/pytest_wide64_p32.log
for f in r2_default wide64 wide64_p32 p32; do cut -c1-120 gpurun_out/bench_$f.json; echo; done
