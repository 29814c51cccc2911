#!/bin/bash
# Round 2, call 2: first run of the CTA-pair GEMM kernel: parity, then the full GPU suite, then bench with and without it.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_tc.py -x -q -k "pair or tc_matches or conj_swap" > gpurun_out/pytest_pair.log 2>&1
echo "pytest pair rc=$?" >> gpurun_out/pytest_pair.log
tail -15 gpurun_out/pytest_pair.log
if grep -q "pytest pair rc=0" gpurun_out/pytest_pair.log; then
  timeout 500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  tail -5 gpurun_out/pytest_gpu.log
  timeout 200 python bench.py --no-cpu --dump-steps gpurun_out/r2_steps_pair.json > gpurun_out/r2_bench_pair.json 2> gpurun_out/r2_bench_pair.err
  cut -c1-300 gpurun_out/r2_bench_pair.json
  TNB_GEMM_PAIR=0 timeout 200 python bench.py --no-cpu --dump-steps gpurun_out/r2_steps_nopair.json > gpurun_out/r2_bench_nopair.json 2> gpurun_out/r2_bench_nopair.err
  cut -c1-300 gpurun_out/r2_bench_nopair.json
  timeout 200 ncu --set full --import-source on --clock-control none -k regex:pair_kernel --launch-skip 1 -c 1 -f -o gpurun_out/r2_gemm_pair \
    python bench.py --steps 1 --warmup 1 --slices-per-step 1 --no-cpu > gpurun_out/ncu_gemm_pair.log 2>&1
fi
