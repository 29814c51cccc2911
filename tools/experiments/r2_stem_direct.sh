#!/bin/bash
# Round 2: direct (register -> global) epilogue of the 1-CTA stem kernel where the planner finds whole 64-byte pieces per
# warp store: parity (incl. bit-identity with the staged path), A/B against TNB_STEM_DIRECT=0 in the same call, planner trace.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_contract.py -m gpu -q -x > gpurun_out/pytest_direct2.log 2>&1; tail -3 gpurun_out/pytest_direct2.log
for v in 1 0 1 0; do
  TNB_STEM_DIRECT=$v timeout 300 python bench.py --no-cpu --no-extras --no-full --dump-steps gpurun_out/r2_steps_direct2_$v.json > gpurun_out/r2_bench_direct2_$v.json 2> gpurun_out/r2_bench_direct2_$v.err
  echo "direct=$v $(cut -c1-120 gpurun_out/r2_bench_direct2_$v.json)"
done
TNB_DEBUG_STEM=1 timeout 200 python bench.py --steps 1 --warmup 1 --slices-per-step 1 --no-cpu --no-extras --no-full 2>&1 >/dev/null | grep "^\[stem\]" | sort | uniq -c > gpurun_out/stem_patterns.txt
cut -c1-230 gpurun_out/stem_patterns.txt
