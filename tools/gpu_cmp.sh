#!/bin/bash
# ours vs the reference cuDNN backend (config 5), GPU parity tests, the bench line.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv | tee gpurun_out/gpu.txt
nproc | tee -a gpurun_out/gpu.txt
echo "== cudnn compare 20bx256"; timeout 900 python tools/cudnn_compare.py --net 20bx256 > gpurun_out/cudnn_compare_20bx256.md 2> gpurun_out/cudnn_compare_20bx256.err; tail -20 gpurun_out/cudnn_compare_20bx256.md; tail -5 gpurun_out/cudnn_compare_20bx256.err
echo "== cudnn compare 10bx128"; timeout 900 python tools/cudnn_compare.py --net 10bx128 --batches 1,8,32,64,256,1024 > gpurun_out/cudnn_compare_10bx128.md 2> gpurun_out/cudnn_compare_10bx128.err; tail -12 gpurun_out/cudnn_compare_10bx128.md
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 2>&1 | tail -1 | cut -c1-2500 | tee gpurun_out/bench.log
