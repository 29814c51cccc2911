#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for t in 0 1; do
for w in mps_norm mps_mpo peps6x6_d4_boundary; do
  TNB_DMMA_TILE=$t timeout 200 python bench.py --workload $w --no-cpu --no-extras > gpurun_out/r2_bench_${w}_t$t.json 2> gpurun_out/r2_bench_${w}_t$t.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_bench_${w}_t$t.json')); r=d['roofline']
print('tile=$t $w', round(d['value'],2), 'TF', round(d['ms_per_step'],3), 'ms', d['gpu_launches'], {n:(round(x['ms'],2),round(x['tflops'],1),round(x['gbs'])) for n,x in r['kernels'].items()})"
done; done
TNB_DMMA_TILE=1 timeout 300 python -m pytest tests -m gpu -x -q -k "c128 or dmma or mps or peps or complex128 or einsum" > gpurun_out/pytest_c128_t1.log 2>&1; tail -3 gpurun_out/pytest_c128_t1.log
