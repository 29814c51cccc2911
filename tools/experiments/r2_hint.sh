#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -k "pair or tc_matches or wide_stem or conj" > gpurun_out/pytest_tc.log 2>&1; tail -3 gpurun_out/pytest_tc.log
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size"
timeout 300 ncu --metrics $M --clock-control none -k regex:"pair_kernel" -c 20 --csv --log-file gpurun_out/r2_ncu_pair_hint.csv \
    python bench.py --steps 1 --warmup 1 --slices-per-step 1 --no-cpu --no-extras --no-full > gpurun_out/ncu_pair_hint.log 2>&1
python tools/ncu_table.py gpurun_out/r2_ncu_pair_hint.csv --min-us 500 2>/dev/null | head -13
timeout 300 python bench.py --no-cpu --no-extras --dump-steps gpurun_out/r2_steps_hint.json > gpurun_out/r2_bench_hint.json 2> gpurun_out/r2_bench_hint.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_hint.json')); r=d['roofline']
print(round(d['value'],2), 'TF', d['clocks']['sm_mhz'], d['full_amplitude']['rel_diff_vs_golden_n1'], {n:(round(x['ms'],2),round(x['tflops'],1),round(x['gbs'])) for n,x in r['kernels'].items()})"
