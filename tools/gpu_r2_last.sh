#!/bin/bash
mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_gpu_parity.py tests/test_gpu_engine.py tests/test_gpu_shapes.py -q -m gpu -x 2>&1 | tail -4 | tee gpurun_out/r02f_pytest_last.log
