#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --durations=5 2>&1 | tail -15 | tee gpurun_out/r2_pytest_gpu.log
timeout 300 python bench.py --steps 50 --warmup 5 2>gpurun_out/r2_bench.err | tail -1 > gpurun_out/r2_bench.json; cut -c1-300 gpurun_out/r2_bench.json
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r2_smoke.log
