#!/bin/bash
# Round 2: latency hiding in the 1-CTA stem kernel (worker prefetch, packed column ranks before the TMEM wait, 8 pairs in
# flight in the write-out): parity, then A/B against TNB_STEM_DEEP=0 in the same call; planner trace of the stem patterns.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_contract.py -m gpu -q -x > gpurun_out/pytest_deep.log 2>&1; tail -3 gpurun_out/pytest_deep.log
for v in 1 0 1 0; do
  TNB_STEM_DEEP=$v timeout 300 python bench.py --no-cpu --no-extras --no-full --dump-steps gpurun_out/r2_steps_deep$v.json > gpurun_out/r2_bench_deep$v.json 2> gpurun_out/r2_bench_deep$v.err
  echo "deep=$v $(cut -c1-120 gpurun_out/r2_bench_deep$v.json)"
done
TNB_DEBUG_STEM=1 timeout 200 python bench.py --steps 1 --warmup 1 --slices-per-step 1 --no-cpu --no-extras --no-full 2>&1 >/dev/null | grep "^\[stem\]" | sort | uniq -c > gpurun_out/stem_patterns.txt
cat gpurun_out/stem_patterns.txt | cut -c1-220
