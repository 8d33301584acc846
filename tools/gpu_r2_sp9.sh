#!/bin/bash
# BASELINE config 1 as self-play with the final library and all host-side replacements: 9x9, 6bx96, 100 visits, finished games
mkdir -p gpurun_out
nproc > gpurun_out/r02f_sp9_cores.txt
python tools/selfplay_bench.py --net 6bx96 --board 9 --playouts 100 --parallel-games 256 --gpus 0 --timeout 100 --label "config1: 9x9, 6bx96, 100 visits, 1 GPU, final library" | tee gpurun_out/r02f_selfplay_9x9.jsonl | cut -c1-600
