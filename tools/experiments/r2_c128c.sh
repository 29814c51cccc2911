#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 400 python -m pytest tests -m gpu -x -q -k "linalg or stem_stream or c128 or dmma or mps or peps or complex128 or graph or einsum or contract" > gpurun_out/pytest_c128.log 2>&1; tail -4 gpurun_out/pytest_c128.log
for w in mps_norm mps_mpo peps6x6_d4_boundary; do
  timeout 200 python bench.py --workload $w --no-cpu --no-extras --dump-steps gpurun_out/r2_steps_$w.json > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_bench_$w.json')); r=d['roofline']
print('$w', round(d['value'],2), 'TF', round(d['ms_per_step'],3), 'ms', d['gpu_launches'], {n:(round(x['ms'],2),round(x['tflops'],1),round(x['gbs'])) for n,x in r['kernels'].items()})"
done
timeout 500 ncu --set full --import-source on --clock-control none -k regex:dmma_kernel --launch-skip 40 -c 2 -f -o gpurun_out/r2_c128_v3 \
    python bench.py --workload mps_mpo --steps 1 --warmup 1 --no-cpu --no-extras > gpurun_out/ncu_c128_v3.log 2>&1
tail -2 gpurun_out/ncu_c128_v3.log
