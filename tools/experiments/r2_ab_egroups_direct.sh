#!/bin/bash
# A/B: epilogue groups of the 1-CTA stem kernel now that eligible steps store directly (no staging tile)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for v in 2 1 2 1; do
  TNB_STEM_EGROUPS=$v timeout 300 python bench.py --no-cpu --no-extras --no-full --dump-steps gpurun_out/r2_steps_eg${v}d.json > gpurun_out/r2_bench_eg${v}d.json 2> gpurun_out/r2_bench_eg${v}d.err
  echo "egroups=$v $(cut -c1-120 gpurun_out/r2_bench_eg${v}d.json)"
done
