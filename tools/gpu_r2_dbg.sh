#!/bin/bash
mkdir -p gpurun_out
python tools/selfplay_bench.py --preset config3 --gpus 0 --parallel-games 64 --window 25 --fp16 --label "dbg" | tee gpurun_out/r2_dbg_window.json | cut -c1-1500
