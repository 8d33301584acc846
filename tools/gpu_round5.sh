#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
echo "== resident weights / residual prefetch"
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/resident.log
import os, sys, tempfile
sys.path.insert(0, '.')
import numpy as np
from sayuri_b200 import engine, synth
for net, b in (("10bx128", 256), ("10bx128", 1024), ("6bx96", 256), ("20bx256", 256)):
    path = os.path.join(tempfile.gettempdir(), "rw_%s.bin" % net); synth.write_synth_net(path, net, seed=20260417)
    pos = synth.synth_positions(64, 19, seed=5).reshape(64, -1)
    for prec in (0, 1):
        pipe = engine.B200ForwardPipe().initialize(path, 19, b, gpus=[0], precision=prec)
        planes = [pos[i % 64] for i in range(b)]
        pipe.batch_forward(0, planes, [19]*b, [0]*b)
        for rw in (0, 1, 0, 1):
            pipe.set_option("resident_weights", rw)
            pipe.time_forward(0, 0, 5, flush_l2=True)
            ms, cms, cn = pipe.time_forward(0, 0, 30, flush_l2=True, profile_conv=True)
            print("%s batch %d precision %d resident %d: %.4f ms median, %.0f evals/s (convs %.3f ms / %d)" % (net, b, prec, rw, float(np.median(ms)), b / float(np.median(ms)) * 1e3, cms, cn), flush=True)
        pipe.destroy()
PY
echo "== conv stats"
for prec in 0 1; do timeout 120 python tools/conv_stats.py --precision $prec --launch 2 2>&1 | grep -E "^net|mma_total|wait_|epi_"; done | tee gpurun_out/stats.log
echo "== bench"; timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1500 | tee gpurun_out/bench.log
echo "== bench fp16"; timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --precision fp16 2>&1 | tail -1 | cut -c1-1500 | tee gpurun_out/bench_fp16.log
