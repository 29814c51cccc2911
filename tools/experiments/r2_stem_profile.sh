#!/bin/bash
# Round 2: single-pass long-K stem routing + source-level profile of the 1-CTA stem kernel and the k-reduction.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -q -x -k "long_k or stem" > gpurun_out/pytest_longk.log 2>&1; tail -3 gpurun_out/pytest_longk.log
timeout 300 python bench.py --no-cpu --no-extras --no-full --dump-steps gpurun_out/r2_steps_longk.json > gpurun_out/r2_bench_longk.json 2> gpurun_out/r2_bench_longk.err
cut -c1-200 gpurun_out/r2_bench_longk.json
timeout 400 ncu --set full --import-source on --clock-control none -k regex:c64_tf32x3_stem_kernel -c 4 -f -o gpurun_out/r2_full_stem1 \
    python bench.py --steps 1 --warmup 1 --slices-per-step 1 --no-cpu --no-extras --no-full > gpurun_out/ncu_full_stem1.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none -k regex:kred -c 1 -f -o gpurun_out/r2_full_kred \
    python bench.py --steps 1 --warmup 1 --slices-per-step 1 --no-cpu --no-extras --no-full > gpurun_out/ncu_full_kred.log 2>&1
du -sh gpurun_out; ls -la gpurun_out/*.ncu-rep
