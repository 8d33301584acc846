#!/bin/bash
# 2 GPUs, final library: the engine tests that need two devices (NCCL broadcast, lanes) and sb_eval from native threads.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02f_2gpu_box.txt; nproc >> gpurun_out/r02f_2gpu_box.txt
timeout 400 python -m pytest tests/test_gpu_engine.py -q -m gpu 2>&1 | tail -4 | tee gpurun_out/r02f_2gpu_pytest_engine.log
timeout 200 python tools/eval_bench.py --gpus 0,1 --threads 1024 --seconds 3 2>&1 | tail -6 | tee gpurun_out/r02f_eval_bench_2gpu.log
timeout 200 python tools/eval_bench.py --gpus 0,1 --threads 1024 --seconds 3 --precision 1 2>&1 | tail -6 | tee -a gpurun_out/r02f_eval_bench_2gpu.log
