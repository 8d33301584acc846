#!/bin/bash
# 4-GPU call: BASELINE config 4 (mixed 9/13/19 boards, 15bx192, Gumbel self-play on 4 GPUs, one process) + sb_eval on 4 replicas
mkdir -p gpurun_out
{ nproc; lscpu | grep -E "Model name|Socket"; nvidia-smi -L | wc -l; } > gpurun_out/r2_4gpu_box.txt 2>&1
ALL=0,1,2,3
{
python tools/eval_bench.py --gpus $ALL --threads 2048 --seconds 2.5
python tools/eval_bench.py --gpus $ALL --threads 32 --async-depth 64 --seconds 2.5 --precision 1
python tools/eval_bench.py --net 15bx192 --gpus $ALL --threads 2048 --seconds 2.5
} 2>&1 | grep -v "^NCCL version" | tee gpurun_out/r2_4gpu_eval_bench.log
python tools/selfplay_bench.py --preset config4 --gpus $ALL --parallel-games 512 --timeout 600 --label "config4: mixed boards, 15bx192, Gumbel, 4 GPUs, one process" | tee gpurun_out/r2_selfplay_4gpu.jsonl | cut -c1-900
python tools/selfplay_bench.py --preset config4 --gpus $ALL --parallel-games 512 --fp16 --timeout 600 --label "config4, front-end default precision (fp16 rung)" | tee -a gpurun_out/r2_selfplay_4gpu.jsonl | cut -c1-900
