#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 400 python -m pytest tests/test_gpu_tc.py -x -q -k "stem" > gpurun_out/pytest_stem.log 2>&1; tail -4 gpurun_out/pytest_stem.log
for e in 2 1; do
  TNB_STEM_EGROUPS=$e timeout 300 python bench.py --no-cpu --no-extras --no-full --dump-steps gpurun_out/r2_steps_eg$e.json > gpurun_out/r2_bench_eg$e.json 2> gpurun_out/r2_bench_eg$e.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_bench_eg$e.json')); r=d['roofline']
print('egroups=$e', round(d['value'],2), 'TF', d['clocks']['sm_mhz'], {n:(round(x['ms'],2),round(x['tflops'],1),round(x['gbs'])) for n,x in r['kernels'].items()})"
done
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
