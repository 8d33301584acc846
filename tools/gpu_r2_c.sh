#!/bin/bash
mkdir -p gpurun_out
rm -f /tmp/sb_fullnets_parity.log
timeout 900 python -m pytest tests/test_gpu_fullnets.py -q 2>&1 | tail -8 > gpurun_out/r2_pytest_c.log
cp /tmp/sb_fullnets_parity.log gpurun_out/r2_fullnets_c.log
