#!/bin/bash
mkdir -p gpurun_out
rm -f /tmp/sb_visit_parity.log
timeout 1200 python -m pytest tests -q -m gpu --durations=5 2>&1 | tail -12 | tee gpurun_out/r02f_pytest_gpu.log
cp /tmp/sb_visit_parity.log gpurun_out/r02f_visit_parity.log 2>/dev/null
