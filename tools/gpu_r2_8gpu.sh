#!/bin/bash
# 8-GPU call (one process drives all GPUs, like the reference front-end): sb_eval scaling through the per-GPU lanes,
# in-engine NCCL weight broadcast over 8 replicas, self-play config 4 settings on 8 GPUs (finished games), config 3 window.
mkdir -p gpurun_out
{ nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)"; nvidia-smi -L | wc -l; } > gpurun_out/r2_8gpu_box.txt 2>&1
ALL=0,1,2,3,4,5,6,7
{
python tools/eval_bench.py --gpus 0 --threads 512 --seconds 2.5
python tools/eval_bench.py --gpus $ALL --threads 2048,4096 --seconds 2.5
python tools/eval_bench.py --gpus 0 --threads 512 --seconds 2.5 --precision 1
python tools/eval_bench.py --gpus $ALL --threads 4096 --seconds 2.5 --precision 1
python tools/eval_bench.py --gpus $ALL --threads 64 --async-depth 64 --seconds 2.5 --precision 1
} 2>&1 | grep -v "^NCCL version" | tee gpurun_out/r2_8gpu_eval_bench.log
python tools/selfplay_bench.py --preset config4 --gpus $ALL --parallel-games 512 --timeout 600 --label "config4 settings, 8 GPUs, one process" | tee gpurun_out/r2_selfplay_8gpu.jsonl | cut -c1-900
python tools/selfplay_bench.py --preset config3 --gpus $ALL --window 70 --fp16 --label "config3 (20bx256, -p 800, 512 games, 8 GPUs), 70 s window, front-end default precision" | tee -a gpurun_out/r2_selfplay_8gpu.jsonl | cut -c1-900
