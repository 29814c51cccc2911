#!/bin/bash
# Round 2 evidence pass (kept under the 64 MiB return limit: CSV tables + a few single-launch reports).
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__registers_per_thread,sm__cycles_elapsed.avg.per_second"
# (1) every launch of the bench command with its device time (shares, not absolutes)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_ncu_launches.csv \
    python bench.py --steps 1 --warmup 1 --slices-per-step 2 --no-cpu --no-extras --no-full > gpurun_out/ncu_launches.log 2>&1
# (2) metric table of every specialised-kernel launch of one slice (+ hoisted steps)
timeout 500 ncu --metrics $M --clock-control none -k regex:"pair_kernel|stem_kernel|kred|acc_kernel" -c 160 --csv --log-file gpurun_out/r2_ncu_c64_kernels.csv \
    python bench.py --steps 1 --warmup 1 --slices-per-step 1 --no-cpu --no-extras --no-full > gpurun_out/ncu_c64_kernels.log 2>&1
timeout 500 ncu --metrics $M --clock-control none -k regex:"dmma_kernel|stem_kernel" --launch-skip 60 -c 40 --csv --log-file gpurun_out/r2_ncu_c128_kernels.csv \
    python bench.py --workload mps_mpo --steps 1 --warmup 1 --no-cpu --no-extras > gpurun_out/ncu_c128_kernels.log 2>&1
# (3) ncu --set full, single launches (source-level stall data)
timeout 300 ncu --set full --import-source on --clock-control none -k regex:pair_kernel --launch-skip 1 -c 4 -f -o gpurun_out/r2_full_pair \
    python bench.py --steps 1 --warmup 1 --slices-per-step 1 --no-cpu --no-extras --no-full > gpurun_out/ncu_full_pair.log 2>&1
timeout 500 ncu --set full --import-source on --clock-control none -k regex:dmma_kernel --launch-skip 60 -c 1 -f -o gpurun_out/r2_full_c128 \
    python bench.py --workload mps_mpo --steps 1 --warmup 1 --no-cpu --no-extras > gpurun_out/ncu_full_c128.log 2>&1
timeout 500 ncu --set full --import-source on --clock-control none -k regex:stem_kernel --launch-skip 30 -c 1 -f -o gpurun_out/r2_full_c128stem \
    python bench.py --workload mps_mpo --steps 1 --warmup 1 --no-cpu --no-extras > gpurun_out/ncu_full_c128stem.log 2>&1
# (4) the default line and the reference arm
timeout 700 python bench.py --dump-steps gpurun_out/r2_steps_default.json > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
echo "bench rc=$?" >> gpurun_out/r2_bench_default.err
cut -c1-200 gpurun_out/r2_bench_default.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
cut -c1-300 gpurun_out/r2_bench_reference.json
du -sh gpurun_out; ls -la gpurun_out/*.ncu-rep
