#!/bin/bash
mkdir -p gpurun_out
echo "== probe full nets"; timeout 600 python tools/gpu_probe.py --shapes 10bx128 --modes split --n 6 2>&1 | cut -c1-330 | tee gpurun_out/probe3.log
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== conv stats split"; timeout 120 python tools/conv_stats.py 2>&1 | tail -12 | tee gpurun_out/stats_split.log
echo "== conv stats fp16"; timeout 120 python tools/conv_stats.py --precision 1 2>&1 | tail -12 | tee gpurun_out/stats_fp16.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1800 | tee gpurun_out/bench.log
echo "== bench fp16"; timeout 300 python bench.py --steps 20 --warmup 5 --precision fp16 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300 | tee gpurun_out/bench_fp16.log
