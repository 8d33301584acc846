#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_engine.py -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02f_pytest_engine_guard.log
