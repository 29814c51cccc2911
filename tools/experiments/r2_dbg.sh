#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 300 python tools/debug_stem.py > gpurun_out/debug_stem.log 2>&1; tail -30 gpurun_out/debug_stem.log
