#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 400 python -m pytest tests/test_gpu_tc.py -x -q > gpurun_out/pytest_tc.log 2>&1
echo "pytest tc rc=$?" >> gpurun_out/pytest_tc.log
tail -12 gpurun_out/pytest_tc.log
timeout 300 python tools/debug_pair.py > gpurun_out/debug_pair.log 2>&1; grep -c "bad 0 of" gpurun_out/debug_pair.log; grep "bad [1-9]" gpurun_out/debug_pair.log
timeout 300 python bench.py --no-cpu --no-extras --dump-steps gpurun_out/r2_steps_staged.json > gpurun_out/r2_bench_staged.json 2> gpurun_out/r2_bench_staged.err
cut -c1-200 gpurun_out/r2_bench_staged.json; tail -3 gpurun_out/r2_bench_staged.err
TNB_STEM_KMAX=512 timeout 300 python bench.py --no-cpu --no-extras --no-full --dump-steps gpurun_out/r2_steps_staged_k512.json > gpurun_out/r2_bench_staged_k512.json 2> gpurun_out/r2_bench_staged_k512.err
cut -c1-200 gpurun_out/r2_bench_staged_k512.json
timeout 500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
