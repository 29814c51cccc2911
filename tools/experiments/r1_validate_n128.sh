#!/bin/bash
# Validation of the new defaults (k-reduction on mma.sync, 128-column stem passes): full GPU suite, default bench line,
# smoke, ncu launch list of one slice, one full ncu capture of the 128-column stem kernel.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -rA --durations=10 > gpurun_out/pytest_gpu_r4.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r4.log
timeout 200 python bench.py --dump-steps gpurun_out/steps_r4.json > gpurun_out/bench_r4.json 2> gpurun_out/bench_r4.err
echo "bench rc=$?" >> gpurun_out/bench_r4.err
timeout 100 python __graft_entry__.py smoke > gpurun_out/smoke_r4.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke_r4.log
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r4.csv \
    python bench.py --steps 1 --warmup 1 --slices-per-step 1 --no-cpu > gpurun_out/ncu_launches_r4.log 2>&1
timeout 150 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k "regex:stem_kernel<128" -c 1 -f -o gpurun_out/r1_stem128 \
    python bench.py --steps 1 --warmup 1 --slices-per-step 1 --no-cpu > gpurun_out/ncu_stem128.log 2>&1
tail -3 gpurun_out/pytest_gpu_r4.log; tail -2 gpurun_out/smoke_r4.log; cut -c1-200 gpurun_out/bench_r4.json; echo; tail -3 gpurun_out/ncu_stem128.log
