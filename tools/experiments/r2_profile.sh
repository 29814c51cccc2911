#!/bin/bash
# Round 2 evidence pass: launch list of the default bench command, ncu --set full of the dominant kernels, default bench line.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
# (1) every launch of the bench command with its device time (shares, not absolutes)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_ncu_launches.csv \
    python bench.py --steps 1 --warmup 1 --slices-per-step 2 --no-cpu --no-extras --no-full > gpurun_out/ncu_launches.log 2>&1
# (2) ncu --set full of the dominant kernels (one slice of the committed path)
timeout 300 ncu --set full --import-source on --clock-control none -k regex:pair_kernel --launch-skip 0 -c 12 -f -o gpurun_out/r2_full_pair \
    python bench.py --steps 1 --warmup 1 --slices-per-step 1 --no-cpu --no-extras --no-full > gpurun_out/ncu_full_pair.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"stem_kernel|kred" --launch-skip 0 -c 14 -f -o gpurun_out/r2_full_stem \
    python bench.py --steps 1 --warmup 1 --slices-per-step 1 --no-cpu --no-extras --no-full > gpurun_out/ncu_full_stem.log 2>&1
timeout 500 ncu --set full --import-source on --clock-control none -k regex:"dmma_kernel" --launch-skip 40 -c 2 -f -o gpurun_out/r2_full_c128 \
    python bench.py --workload mps_mpo --steps 1 --warmup 1 --no-cpu --no-extras > gpurun_out/ncu_full_c128.log 2>&1
timeout 500 ncu --set full --import-source on --clock-control none -k regex:"stem_kernel" --launch-skip 30 -c 1 -f -o gpurun_out/r2_full_c128stem \
    python bench.py --workload mps_mpo --steps 1 --warmup 1 --no-cpu --no-extras > gpurun_out/ncu_full_c128stem.log 2>&1
# (3) tile = 2 experiment for the FP64 kernel
for w in mps_mpo peps6x6_d4_boundary; do
  TNB_DMMA_TILE=2 timeout 200 python bench.py --workload $w --no-cpu --no-extras > gpurun_out/r2_bench_${w}_t2.json 2> gpurun_out/r2_bench_${w}_t2.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_bench_${w}_t2.json')); r=d['roofline']
print('tile=2 $w', round(d['value'],2), 'TF', round(d['ms_per_step'],3), 'ms', {n:(round(x['ms'],2),round(x['tflops'],1),round(x['gbs'])) for n,x in r['kernels'].items()})"
done
# (4) the default line
timeout 700 python bench.py --dump-steps gpurun_out/r2_steps_default.json > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err
echo "bench rc=$?" >> gpurun_out/r2_bench_default.err
cut -c1-200 gpurun_out/r2_bench_default.json
ls -la gpurun_out/*.ncu-rep
