#!/bin/bash
# 2 GPUs: hardware multi-GPU parity tests + bench at N = 2 (weak-scaling steps + full-amplitude strong scaling)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
nvidia-smi -L > gpurun_out/smi_L.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -rA > gpurun_out/pytest_multi.log 2>&1
echo "pytest multi rc=$?" >> gpurun_out/pytest_multi.log
tail -15 gpurun_out/pytest_multi.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "bench n2 rc=$?" >> gpurun_out/r2_bench_n2.err
grep '^{' gpurun_out/r2_bench_n2.json | cut -c1-300; tail -3 gpurun_out/r2_bench_n2.err
