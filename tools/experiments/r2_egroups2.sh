#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 400 python -m pytest tests/test_gpu_tc.py -q -k "stem" > gpurun_out/pytest_stem2.log 2>&1; grep -E "FAILED|passed|failed" gpurun_out/pytest_stem2.log | head -30
TNB_STEM_EGROUPS=1 timeout 400 python -m pytest tests/test_gpu_tc.py -q -k "stem" > gpurun_out/pytest_stem1.log 2>&1; grep -E "FAILED|passed|failed" gpurun_out/pytest_stem1.log | head
