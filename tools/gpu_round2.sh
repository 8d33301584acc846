#!/bin/bash
# round-1 session-2 checks: parity tests, conv ablations (ring depth, tail split), batcher throughput, front-end
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== conv stats: default build (NB=8, NA=3)"
for prec in 0 1; do for ts in 1 0; do
  timeout 120 python tools/conv_stats.py --precision $prec --launch 2 --tail-split $ts 2>&1 | grep -E "^net|mma_total|wait_b|epi_total"
done; done | tee gpurun_out/stats_default.log
echo "== conv stats: NB=16 NA=2 build"
for prec in 0 1; do
  SAYURI_B200_LIB=$PWD/sayuri_b200/libsayuri_b200_nb16.so timeout 120 python tools/conv_stats.py --precision $prec --launch 2 2>&1 | grep -E "^net|mma_total|wait_b|wait_slab|epi_total"
done | tee gpurun_out/stats_nb16.log
echo "== bench"; timeout 600 python bench.py --steps 30 --warmup 5 2>&1 | tail -1 | cut -c1-2600 | tee gpurun_out/bench.log
echo "== eval bench (sb_eval, native threads)"
timeout 300 python tools/eval_bench.py --net 10bx128 --threads 16,64,256,512,1024 --seconds 2 2>&1 | tee gpurun_out/eval_bench.log
timeout 300 python tools/eval_bench.py --net 10bx128 --threads 512 --seconds 2 --precision 1 2>&1 | tee -a gpurun_out/eval_bench.log
echo "== blocking sb_forward_batch through pageable host buffers"
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/host_path.log
import os, sys, tempfile
sys.path.insert(0, '.')
import numpy as np
from sayuri_b200 import engine, synth
path = os.path.join(tempfile.gettempdir(), "hp_10bx128.bin"); synth.write_synth_net(path, "10bx128", seed=20260417)
pos = synth.synth_positions(64, 19, seed=5).reshape(64, -1)
for b in (32, 256, 1024):
    pipe = engine.B200ForwardPipe().initialize(path, 19, b, gpus=[0])
    planes = [pos[i % 64] for i in range(b)]
    for pk in (0, 1):
        pipe.set_option("pack_inputs", pk)
        pipe.batch_forward(0, planes, [19]*b, [0]*b)
        ev, ms = pipe.time_batch_forward_host(0, planes, [19]*b, [0]*b, 1.5)
        print("sb_forward_batch batch %d pack_inputs %d: %.0f evals/s (%.3f ms/call)" % (b, pk, ev, ms), flush=True)
    pipe.destroy()
PY
echo "== front-end netbench: engine batcher vs reference batcher"
W=/tmp/fe_10bx128.bin
python -c "
import sys; sys.path.insert(0,'.')
from sayuri_b200 import synth
synth.write_synth_net('$W', '10bx128', seed=20260417)"
FE=oracle/_ref/sayuri_b200_frontend
printf 'netbench timelimit 5 batchsize 256\nquit\n' | timeout 200 $FE -w $W --no-fp16 -g 0 -b 256 2>&1 | grep -E "batch size=|rror" | sed 's/^/engine-batcher: /' | tee gpurun_out/netbench.log
printf 'netbench timelimit 5 batchsize 256\nquit\n' | SAYURI_B200_REF_BATCHER=1 timeout 200 $FE -w $W --no-fp16 -g 0 -b 256 2>&1 | grep -E "batch size=|rror" | sed 's/^/reference-batcher: /' | tee -a gpurun_out/netbench.log
echo "== visit parity 9x9 (engine batcher in the det front-end)"
timeout 600 python tools/visit_parity.py --net 6bx96 --board 9 --playouts 400 --moves 4 --seeds 1,2 2>&1 | tail -2 | tee gpurun_out/visit_parity.log
