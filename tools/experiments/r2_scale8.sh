#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
NG=$(nvidia-smi -L | wc -l); echo "gpus: $NG"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/r2_bench_n$NG.json 2> gpurun_out/r2_bench_n$NG.err
echo "rc=$?"; grep '^{' gpurun_out/r2_bench_n$NG.json | cut -c1-250; tail -2 gpurun_out/r2_bench_n$NG.err | cut -c1-200
