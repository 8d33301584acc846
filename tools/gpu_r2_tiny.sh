#!/bin/bash
mkdir -p gpurun_out
timeout 18 python -m pytest tests/test_gpu_frontend.py -q -m gpu -k 13x13 2>&1 | tail -3 | tee gpurun_out/r02f_pytest_frontend_13.log
