#!/bin/bash
# 2 GPUs, final sources: multi-GPU parity tests, the stem_tc tests (incl. the new direct-epilogue layouts), bench at N = 2
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -rA > gpurun_out/pytest_multi.log 2>&1
echo "pytest multi rc=$?" >> gpurun_out/pytest_multi.log
tail -6 gpurun_out/pytest_multi.log
timeout 400 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "stem" > gpurun_out/pytest_stem_final.log 2>&1; tail -2 gpurun_out/pytest_stem_final.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "bench n2 rc=$?" >> gpurun_out/r2_bench_n2.err
grep '^{' gpurun_out/r2_bench_n2.json | cut -c1-300; tail -3 gpurun_out/r2_bench_n2.err
