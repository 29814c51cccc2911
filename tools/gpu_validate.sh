#!/bin/bash
# Final validation of the committed binary + ncu --set full captures of the stem kernels (evidence for the next round).
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -q -rA --durations=5 > gpurun_out/pytest_gpu_r6.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r6.log
timeout 200 python bench.py --dump-steps gpurun_out/steps_r6.json > gpurun_out/bench_r6.json 2> gpurun_out/bench_r6.err
echo "bench rc=$?" >> gpurun_out/bench_r6.err
timeout 100 python __graft_entry__.py smoke > gpurun_out/smoke_r6.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke_r6.log
timeout 150 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:stem_kernel<\(int\)(64|32)' -c 5 -f -o gpurun_out/r1_stem_small \
    python bench.py --steps 1 --warmup 1 --slices-per-step 1 --no-cpu > gpurun_out/ncu_stem_small.log 2>&1
timeout 150 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k 'regex:stem_kernel<\(int\)128' --launch-skip 16 -c 2 -f -o gpurun_out/r1_stem128 \
    python bench.py --steps 1 --warmup 1 --slices-per-step 1 --no-cpu > gpurun_out/ncu_stem128.log 2>&1
tail -3 gpurun_out/pytest_gpu_r6.log; tail -2 gpurun_out/smoke_r6.log; cut -c1-160 gpurun_out/bench_r6.json; echo; tail -3 gpurun_out/ncu_stem_small.log | cut -c1-200; tail -3 gpurun_out/ncu_stem128.log | cut -c1-200
