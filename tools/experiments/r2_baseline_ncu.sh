#!/bin/bash
# Round 2, call 1: baseline of the round-1 code on this round's box + source-level ncu captures of the kernels the
# verdict names (tile GEMM, c128 DMMA, c128 SIMT stem).   gpurun --timeout 900 -- 'bash tools/experiments/r2_baseline_ncu.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,memory.total --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 200 python bench.py --no-cpu --dump-steps gpurun_out/r2_steps_base.json > gpurun_out/r2_bench_base.json 2> gpurun_out/r2_bench_base.err
echo "bench rc=$?" >> gpurun_out/r2_bench_base.err
timeout 200 ncu --set full --import-source on --clock-control none -k regex:acc_kernel --launch-skip 2 -c 2 -f -o gpurun_out/r2_gemm_base \
    python bench.py --steps 1 --warmup 1 --slices-per-step 1 --no-cpu > gpurun_out/ncu_gemm_base.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:dmma --launch-skip 60 -c 2 -f -o gpurun_out/r2_c128_base \
    python bench.py --workload mps_mpo --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_c128_base.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:stem_kernel --launch-skip 30 -c 1 -f -o gpurun_out/r2_c128stem_base \
    python bench.py --workload mps_mpo --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_c128stem_base.log 2>&1
timeout 120 python bench.py --workload mps_mpo --no-cpu --dump-steps gpurun_out/r2_steps_mpsmpo_base.json > gpurun_out/r2_bench_mpsmpo_base.json 2> gpurun_out/r2_bench_mpsmpo_base.err
cut -c1-400 gpurun_out/r2_bench_base.json; echo; tail -2 gpurun_out/ncu_gemm_base.log | cut -c1-200; tail -2 gpurun_out/ncu_c128_base.log | cut -c1-200; tail -2 gpurun_out/ncu_c128stem_base.log | cut -c1-200
ls -la gpurun_out
