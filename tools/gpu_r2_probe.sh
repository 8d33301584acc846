#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/precision_probe.py 2>&1 | tee gpurun_out/r2_precision_probe2.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_gpu_engine.py -q 2>&1 | tail -30 > gpurun_out/r2_pytest_b.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2_bench_b.err | tail -1 > gpurun_out/r2_bench_b.json
