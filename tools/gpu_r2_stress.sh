#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/determinism_stress.py --repeats 1200 2>&1 | tee gpurun_out/r2_stress.log
SAYURI_B200_OPTIONS=conv_chain=0,layer_overlap=0 timeout 300 python -m pytest tests/test_gpu_frontend.py -q -m gpu -x -k identical_root 2>&1 | tail -4 | tee gpurun_out/r2_frontend_plain.log
timeout 300 python -m pytest tests/test_gpu_frontend.py -q -m gpu -x -k identical_root 2>&1 | grep -E "passed|failed|different move|AssertionError:" | tail -4 | tee gpurun_out/r2_frontend_default.log
