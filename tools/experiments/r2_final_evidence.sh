#!/bin/bash
# Round 2, final state: ncu launch list + kernel metric tables of the committed sources, traffic json regenerated from them
# (stamped with the source digest), then the full GPU suite, smoke, the default bench line and the reference arm.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__registers_per_thread,sm__cycles_elapsed.avg.per_second"
# (1) every launch of the bench command with its device time (shares, not absolutes)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_ncu_launches.csv \
    python bench.py --steps 1 --warmup 1 --slices-per-step 2 --no-cpu --no-extras --no-full > gpurun_out/ncu_launches.log 2>&1
# (2) metric table of every specialised-kernel launch of one slice (+ hoisted steps)
timeout 500 ncu --metrics $M --clock-control none -k regex:"pair_kernel|stem_kernel|kred|acc_kernel" -c 160 --csv --log-file gpurun_out/r2_ncu_c64_kernels.csv \
    python bench.py --steps 1 --warmup 1 --slices-per-step 1 --no-cpu --no-extras --no-full > gpurun_out/ncu_c64_kernels.log 2>&1
timeout 500 ncu --metrics $M --clock-control none -k regex:"dmma_kernel|stem_kernel" --launch-skip 60 -c 40 --csv --log-file gpurun_out/r2_ncu_c128_kernels.csv \
    python bench.py --workload mps_mpo --steps 1 --warmup 1 --no-cpu --no-extras > gpurun_out/ncu_c128_kernels.log 2>&1
# (3) DRAM traffic of the dominant launches vs their algorithmic bytes, stamped with the digest of these sources
python tools/traffic_from_csv.py \
    "sycamore53_m14:c64_tf32x3:gpurun_out/r2_ncu_c64_kernels.csv:pair_kernel<\(bool\)0>|pair_kernel<0>:512:2097152:256:8" \
    "sycamore53_m14:stem_tc:gpurun_out/r2_ncu_c64_kernels.csv:pair_kernel<\(bool\)1>|pair_kernel<1>:128:8388608:128:8" \
    "mps_mpo:c128_dmma:gpurun_out/r2_ncu_c128_kernels.csv:dmma_kernel:3072:2048:1024:16" --out gpurun_out/r2_traffic.json > gpurun_out/traffic.log 2>&1
cp gpurun_out/r2_traffic.json profiles/r2_traffic.json
grep -E '"ratio"|"launch"' gpurun_out/r2_traffic.json
# (4) the full GPU suite, smoke, the default line, the reference arm
timeout 800 python -m pytest tests -m gpu -q -rA --durations=8 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -3; grep -E "^FAILED|^ERROR" gpurun_out/pytest_gpu.log | head
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 700 python bench.py --dump-steps gpurun_out/r2_steps_default.json > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
echo "bench rc=$?" >> gpurun_out/r2_bench_default.err
cut -c1-300 gpurun_out/r2_bench_default.json; tail -3 gpurun_out/r2_bench_default.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
cut -c1-300 gpurun_out/r2_bench_reference.json
# (5) SIMT stem kernel: direct vs staged form on rows-fastest outputs (one process each)
for v in 1 0 1 0; do TNB_STEM_SIMT_DIRECT=$v timeout 120 python tools/probes/simt_direct_probe.py 2>&1 | tail -1; done > gpurun_out/simt_direct_probe.txt
cat gpurun_out/simt_direct_probe.txt | cut -c1-600
# (6) experiment: all-columns direct rule for the SIMT form (TNB_STEM_DIRECT=3) on the skinny MPO step of configs[4]
for v in 3 2; do
  TNB_STEM_DIRECT=$v timeout 300 python bench.py --workload mps_mpo --no-cpu --no-extras > gpurun_out/r2_bench_mpo_mode$v.json 2> gpurun_out/r2_bench_mpo_mode$v.err
  echo "mps_mpo TNB_STEM_DIRECT=$v $(python -c "import json;d=json.load(open('gpurun_out/r2_bench_mpo_mode$v.json'));print(d['value'], d['roofline']['kernels']['stem'])")"
done
du -sh gpurun_out
