#!/bin/bash
# Round 2, first GPU call: the new parity tests (full-depth nets, visit counts through the shim, engine behaviour), then
# the bench line of both arms.
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/r2_gpu.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -q -m gpu --durations=8 2>&1 | tail -25 | tee gpurun_out/r2_pytest_gpu.log
cp /tmp/sb_fullnets_parity.log /tmp/sb_visit_parity.log /tmp/sb_weights_broadcast.log gpurun_out/ 2>/dev/null
echo "== bench"
timeout 400 python bench.py --steps 50 --warmup 5 2>gpurun_out/r2_bench.err | tail -1 > gpurun_out/r2_bench.json; cut -c1-400 gpurun_out/r2_bench.json
