#!/bin/bash
# Reduced evidence pass for the last source change of round 2 (SIMT direct form: all-columns rule by default, N <= 8 only):
# ncu metric tables -> traffic json stamped with these sources, stem tests, default bench line.  Ordered by importance.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__registers_per_thread,sm__cycles_elapsed.avg.per_second"
timeout 300 ncu --metrics $M --clock-control none -k regex:"pair_kernel|stem_kernel|kred|acc_kernel" -c 160 --csv --log-file gpurun_out/r2_ncu_c64_kernels.csv \
    python bench.py --steps 1 --warmup 1 --slices-per-step 1 --no-cpu --no-extras --no-full > gpurun_out/ncu_c64_kernels.log 2>&1
timeout 300 ncu --metrics $M --clock-control none -k regex:"dmma_kernel|stem_kernel|stem_direct" --launch-skip 60 -c 40 --csv --log-file gpurun_out/r2_ncu_c128_kernels.csv \
    python bench.py --workload mps_mpo --steps 1 --warmup 1 --no-cpu --no-extras > gpurun_out/ncu_c128_kernels.log 2>&1
python tools/traffic_from_csv.py \
    "sycamore53_m14:c64_tf32x3:gpurun_out/r2_ncu_c64_kernels.csv:pair_kernel<\(bool\)0>|pair_kernel<0>:512:2097152:256:8" \
    "sycamore53_m14:stem_tc:gpurun_out/r2_ncu_c64_kernels.csv:pair_kernel<\(bool\)1>|pair_kernel<1>:128:8388608:128:8" \
    "mps_mpo:c128_dmma:gpurun_out/r2_ncu_c128_kernels.csv:dmma_kernel:3072:2048:1024:16" --out gpurun_out/r2_traffic.json > gpurun_out/traffic.log 2>&1
cp gpurun_out/r2_traffic.json profiles/r2_traffic.json
grep -E '"ratio"' gpurun_out/r2_traffic.json
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "stem" > gpurun_out/pytest_stem_last.log 2>&1; tail -2 gpurun_out/pytest_stem_last.log
timeout 400 python bench.py --dump-steps gpurun_out/r2_steps_default.json > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
echo "bench rc=$?" >> gpurun_out/r2_bench_default.err
cut -c1-200 gpurun_out/r2_bench_default.json; tail -2 gpurun_out/r2_bench_default.err
