#!/bin/bash
# 1-GPU self-play measurements (finished games) + the relaxed visit-parity test
mkdir -p gpurun_out
nproc > gpurun_out/r2_sp1_cores.txt
timeout 400 python -m pytest tests/test_gpu_frontend.py -q -k 19x19 2>&1 | tail -3 | tee gpurun_out/r2_visit19b.log
cp /tmp/sb_visit_parity.log gpurun_out/r2_visit_parity_b.log 2>/dev/null
: > gpurun_out/r2_selfplay_1gpu.jsonl
python tools/selfplay_bench.py --preset config4 --gpus 0 --parallel-games 128 --timeout 300 --label "config4 on 1 GPU" | tee -a gpurun_out/r2_selfplay_1gpu.jsonl | cut -c1-700
python tools/selfplay_bench.py --preset config2 --gpus 0 --parallel-games ${1:-128} --timeout 900 --label "config2 self-play, 1 GPU" | tee -a gpurun_out/r2_selfplay_1gpu.jsonl | cut -c1-700
