#!/bin/bash
# 2-GPU checks: (1) the bench contract under torchrun (one process per GPU, NCCL weight broadcast, weak scaling),
# (2) ONE process driving both GPUs through the engine batcher (shared batch ring, 2 workers per GPU).
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/multi_gpus.txt
echo "== torchrun bench --gpus 2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --eval-threads 256 2>&1 | tail -1 | cut -c1-1500 | tee gpurun_out/bench_2gpu.log
echo "== one process, two replicas, sb_eval"
timeout 300 python tools/eval_bench.py --net 10bx128 --threads 256,1024 --seconds 3 2>&1 | tee gpurun_out/eval_bench_2gpu.log
timeout 300 python tools/eval_bench.py --net 10bx128 --threads 1024 --seconds 3 --precision 1 2>&1 | tee -a gpurun_out/eval_bench_2gpu.log
echo "== parity on both replicas of one engine"
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/two_replicas.log
import os, sys
sys.path.insert(0, '.')
import numpy as np
from sayuri_b200 import engine, synth
path = "tests/golden/ref_3bx32.bin.txt"
pipe = engine.B200ForwardPipe().initialize(path, 19, 8, gpus=[0, 1])
assert pipe.get_num_workers() == 2
planes = [synth.synth_positions(1, bs, seed=40 + i)[0].ravel() for i, bs in enumerate((19, 13, 9, 19))]
a = pipe.batch_forward(0, planes, [19, 13, 9, 19], [0, 1, 2, 3])
b = pipe.batch_forward(1, planes, [19, 13, 9, 19], [0, 1, 2, 3])
same = all(np.array_equal(a[f], b[f]) for f in ("probabilities", "ownership", "wdl", "pass_probability"))
print("replica 0 and replica 1 bit-identical:", same, " checksums equal:", pipe.weights_checksum(0) == pipe.weights_checksum(1))
pipe.destroy()
PY
