#!/bin/bash
# Round 2: direct form of the SIMT stem kernel (complex128 skinny step of configs[4]) + relaxed criterion experiment
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_einsum.py -m gpu -q -x > gpurun_out/pytest_simt_direct.log 2>&1; tail -3 gpurun_out/pytest_simt_direct.log
for v in 1 0; do
  TNB_STEM_DIRECT=$v timeout 300 python bench.py --workload mps_mpo --no-cpu --no-extras --dump-steps gpurun_out/r2_steps_mpo_d$v.json > gpurun_out/r2_bench_mpo_d$v.json 2> gpurun_out/r2_bench_mpo_d$v.err
  echo "mps_mpo direct=$v $(cut -c1-140 gpurun_out/r2_bench_mpo_d$v.json)"
done
for v in 2 1 2 1; do
  TNB_STEM_DIRECT=$v timeout 300 python bench.py --no-cpu --no-extras --no-full --dump-steps gpurun_out/r2_steps_relax$v.json > gpurun_out/r2_bench_relax$v.json 2> gpurun_out/r2_bench_relax$v.err
  echo "sycamore direct=$v $(cut -c1-120 gpurun_out/r2_bench_relax$v.json)"
done
