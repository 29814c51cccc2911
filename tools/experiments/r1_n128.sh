#!/bin/bash
# Experiment round: 128-column passes of the stem kernel (TNB_STEM_N128=1), alone and with up to 16 passes.
mkdir -p gpurun_out
TNB_STEM_N128=1 timeout 200 python -m pytest tests -m gpu -q -rA -k "stem_tc or sycamore53 or regular3" > gpurun_out/pytest_n128.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_n128.log
TNB_STEM_N128=1 timeout 150 python bench.py --no-cpu --dump-steps gpurun_out/steps_n128.json > gpurun_out/bench_n128.json 2> gpurun_out/bench_n128.err
TNB_STEM_N128=1 TNB_STEM_MAX_PASSES=16 timeout 150 python -m pytest tests -m gpu -q -rA -k "sycamore53" > gpurun_out/pytest_n128_p16.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_n128_p16.log
TNB_STEM_N128=1 TNB_STEM_MAX_PASSES=16 timeout 150 python bench.py --no-cpu --dump-steps gpurun_out/steps_n128_p16.json > gpurun_out/bench_n128_p16.json 2> gpurun_out/bench_n128_p16.err
tail -4 gpurun_out/pytest_n128.log; tail -3 gpurun_out/pytest_n128_p16.log
for f in n128 n128_p16; do cut -c1-120 gpurun_out/bench_$f.json; echo; tail -2 gpurun_out/bench_$f.err; done
