#!/bin/bash
# 4-GPU call with the ladder-map replacement in the front-end: BASELINE config 4, then config 2 as self-play (finished games)
mkdir -p gpurun_out
nproc > gpurun_out/r2_4gpu_b_cores.txt
ALL=0,1,2,3
python tools/selfplay_bench.py --preset config4 --gpus $ALL --parallel-games 512 --timeout 600 --label "config4: mixed boards, 15bx192, Gumbel, 4 GPUs, one process, + ladder-map replacement" | tee gpurun_out/r2_selfplay_4gpu_ladder.jsonl | cut -c1-700
python tools/selfplay_bench.py --preset config2 --gpus $ALL --parallel-games 256 --timeout 900 --label "config2 self-play, 4 GPUs, one process, + ladder-map replacement" | tee -a gpurun_out/r2_selfplay_4gpu_ladder.jsonl | cut -c1-700
