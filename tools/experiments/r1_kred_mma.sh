#!/bin/bash
# One gpurun call: GPU test suite, default bench line, the TNB_KRED_MMA=1 variant (tests + bench), one more workload,
# smoke, full ncu captures of both k-reduction kernels.  Everything lands in gpurun_out/.
#   gpurun --timeout 780 -- 'bash tools/experiments/r1_kred_mma.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 400 python -m pytest tests -m gpu -q -rA --durations=20 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 200 python bench.py --dump-steps gpurun_out/steps_default.json > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?" >> gpurun_out/bench_n1.err
TNB_KRED_MMA=1 timeout 200 python -m pytest tests -m gpu -q -rA -k "k_reduction or sycamore53 or regular3" > gpurun_out/pytest_gpu_kredmma.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu_kredmma.log
TNB_KRED_MMA=1 timeout 150 python bench.py --no-cpu --dump-steps gpurun_out/steps_kredmma.json > gpurun_out/bench_n1_kredmma.json 2> gpurun_out/bench_n1_kredmma.err
timeout 100 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 100 python bench.py --workload regular3_n100_d4 --no-cpu > gpurun_out/bench_regular3_n100_d4.json 2> gpurun_out/bench_regular3.err
TNB_KRED_MMA=1 timeout 120 ncu --set full --import-source on --clock-control none -k regex:einsum_kred -c 1 -f -o gpurun_out/r1_kred_mma \
    python bench.py --steps 1 --warmup 1 --slices-per-step 1 --no-cpu > gpurun_out/ncu_kred_mma.log 2>&1
timeout 120 ncu --set full --import-source on --clock-control none -k regex:einsum_kred -c 1 -f -o gpurun_out/r1_kred_fma \
    python bench.py --steps 1 --warmup 1 --slices-per-step 1 --no-cpu > gpurun_out/ncu_kred_fma.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu_kredmma.log; tail -2 gpurun_out/smoke.log
cut -c1-600 gpurun_out/bench_n1.json; echo; cut -c1-300 gpurun_out/bench_n1_kredmma.json
