#!/bin/bash
# 1-GPU self-play with the ladder-map replacement linked into the front-end (same command as profiles/r02_selfplay_1gpu.jsonl)
mkdir -p gpurun_out
python tools/selfplay_bench.py --preset config2 --gpus 0 --parallel-games 128 --timeout 900 --label "config2 self-play, 1 GPU, + ladder-map replacement" | tee gpurun_out/r2_selfplay_1gpu_ladder.jsonl | cut -c1-800
