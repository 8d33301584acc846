#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "$N GPUs, $(nproc) host cores" | tee gpurun_out/multi4.txt
echo "== torchrun bench --gpus $N"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 --eval-threads 256 2>&1 | tail -1 | cut -c1-1300 | tee gpurun_out/bench_${N}gpu.log
echo "== torchrun bench --impl reference --gpus $N (rank 0 only)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>&1 | grep impl | cut -c1-300 | tee gpurun_out/bench_ref_${N}gpu.log
echo "== one process, $N replicas, sb_eval"
timeout 300 python tools/eval_bench.py --net 10bx128 --threads 2048 --seconds 3 2>&1 | tee gpurun_out/eval_bench_${N}gpu.log
timeout 300 python tools/eval_bench.py --net 10bx128 --threads 2048 --seconds 3 --precision 1 2>&1 | tee -a gpurun_out/eval_bench_${N}gpu.log
