#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python tools/debug_stem.py > gpurun_out/debug_stem.log 2>&1; tail -6 gpurun_out/debug_stem.log | cut -c1-200
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu --no-extras --dump-steps gpurun_out/r2_steps_eg2b.json > gpurun_out/r2_bench_eg2b.json 2> gpurun_out/r2_bench_eg2b.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_eg2b.json')); r=d['roofline']
print(round(d['value'],2), 'TF', d['clocks']['sm_mhz'], d['full_amplitude'], {n:(round(x['ms'],2),round(x['tflops'],1),round(x['gbs'])) for n,x in r['kernels'].items()})"
