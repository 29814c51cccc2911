#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python tools/debug_pair.py > gpurun_out/debug_pair.log 2>&1
cat gpurun_out/debug_pair.log | tail -40
timeout 200 python bench.py --no-cpu --dump-steps gpurun_out/r2_steps_pair2.json > gpurun_out/r2_bench_pair2.json 2> gpurun_out/r2_bench_pair2.err
cut -c1-200 gpurun_out/r2_bench_pair2.json
timeout 200 ncu --set full --import-source on --clock-control none -k regex:pair_kernel --launch-skip 1 -c 1 -f -o gpurun_out/r2_gemm_pair2 \
    python bench.py --steps 1 --warmup 1 --slices-per-step 1 --no-cpu > gpurun_out/ncu_gemm_pair2.log 2>&1
timeout 200 python bench.py --workload mps_mpo --no-cpu --dump-steps gpurun_out/r2_steps_mpsmpo2.json > gpurun_out/r2_bench_mpsmpo2.json 2> gpurun_out/r2_bench_mpsmpo2.err
cut -c1-200 gpurun_out/r2_bench_mpsmpo2.json
timeout 200 python bench.py --workload peps6x6_d4_boundary --no-cpu > gpurun_out/r2_bench_peps2.json 2> gpurun_out/r2_bench_peps2.err
cut -c1-200 gpurun_out/r2_bench_peps2.json
timeout 200 python -m pytest tests -m gpu -q -k "c128 or dmma or mps or peps or complex128" > gpurun_out/pytest_c128.log 2>&1
tail -3 gpurun_out/pytest_c128.log
