#!/bin/bash
# SIMT stem kernel, direct form with all K loads in flight: A/B on configs[4]; stem tests on the new defaults
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for v in 1 0; do
  TNB_STEM_SIMT_DIRECT=$v timeout 300 python bench.py --workload mps_mpo --no-cpu --no-extras > gpurun_out/r2_bench_mpo_sd$v.json 2> gpurun_out/r2_bench_mpo_sd$v.err
  echo "mps_mpo simt_direct=$v $(python -c "import json;d=json.load(open('gpurun_out/r2_bench_mpo_sd$v.json'));print(d['value'], d['roofline']['kernels']['stem'])")"
done
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "stem" > gpurun_out/pytest_simt_direct2.log 2>&1; tail -3 gpurun_out/pytest_simt_direct2.log
