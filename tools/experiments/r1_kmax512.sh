#!/bin/bash
# Experiment: K <= 512 in the 128-column stem passes (TNB_STEM_KMAX=512): TMEM chunks of 128 k added in the staging tile.
mkdir -p gpurun_out
TNB_STEM_KMAX=512 timeout 200 python -m pytest tests -m gpu -q -rA -k "stem_tc or sycamore53 or regular3" > gpurun_out/pytest_kmax.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_kmax.log
TNB_STEM_KMAX=512 timeout 150 python bench.py --no-cpu --dump-steps gpurun_out/steps_kmax.json > gpurun_out/bench_kmax.json 2> gpurun_out/bench_kmax.err
timeout 100 python -m pytest tests -m gpu -q -rA -k "stem_tc_long_k" > gpurun_out/pytest_longk_default.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_longk_default.log
tail -4 gpurun_out/pytest_kmax.log; tail -3 gpurun_out/pytest_longk_default.log
cut -c1-120 gpurun_out/bench_kmax.json; echo; tail -2 gpurun_out/bench_kmax.err
