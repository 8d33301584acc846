#!/bin/bash
mkdir -p gpurun_out
python tools/selfplay_bench.py --net 6bx96 --board 9 --playouts 100 --parallel-games 512 --gpus 0 --timeout 80 --label "config1: 9x9, 6bx96, 100 visits, 1 GPU, 512 games, final library" | tee gpurun_out/r02f_selfplay_9x9_512.jsonl | cut -c1-600
