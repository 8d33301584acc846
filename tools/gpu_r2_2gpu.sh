#!/bin/bash
# 2-GPU call: in-engine NCCL weight broadcast test, sb_eval in one process on 1 vs 2 replicas, self-play on 2 GPUs
mkdir -p gpurun_out
{ nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)"; nvidia-smi -L; } > gpurun_out/r2_2gpu_box.txt 2>&1
rm -f /tmp/sb_weights_broadcast.log
timeout 300 python -m pytest tests/test_gpu_engine.py -q 2>&1 | tail -4 | tee gpurun_out/r2_2gpu_pytest_engine.log
cp /tmp/sb_weights_broadcast.log gpurun_out/r2_2gpu_weights_broadcast.log 2>/dev/null
{
python tools/eval_bench.py --gpus 0 --threads 512 --seconds 3
python tools/eval_bench.py --gpus 0,1 --threads 512,1024 --seconds 3
python tools/eval_bench.py --gpus 0 --threads 512 --seconds 3 --precision 1
python tools/eval_bench.py --gpus 0,1 --threads 1024 --seconds 3 --precision 1
python tools/eval_bench.py --gpus 0,1 --threads 16 --async-depth 64 --seconds 3 --precision 1
} 2>&1 | tee gpurun_out/r2_2gpu_eval_bench.log
python tools/selfplay_bench.py --preset config2 --gpus 0,1 --parallel-games 256 --timeout 900 --label "config2 self-play, 2 GPUs, one process" | tee gpurun_out/r2_selfplay_2gpu.jsonl | cut -c1-800
