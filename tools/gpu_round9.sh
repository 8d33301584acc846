#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300 | tee gpurun_out/bench.log
echo "== bench fp16"; timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --precision fp16 2>&1 | tail -1 | cut -c1-1800 | tee gpurun_out/bench_fp16.log
echo "== conv stats fp16"; timeout 120 python tools/conv_stats.py --precision 1 --launch 2 2>&1 | grep -E "^net|mma_total|wait_|epi_" | tee gpurun_out/stats_fp16.log
echo "== families"
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/family_speed.log
import os, sys, tempfile
sys.path.insert(0, '.')
import numpy as np
from sayuri_b200 import engine, synth
pos = synth.synth_positions(64, 19, seed=5).reshape(64, -1)
b = 256
for name, base, head, shape in (("10b x128 ResidualBlock", "ResidualBlock", "Normal", (10, 128, 24, 24)), ("10b x128 MixerBlock + RepLK head", "MixerBlock", "RepLK", (10, 128, 24, 24)),
                                ("20b x256 ResidualBlock", "ResidualBlock", "Normal", (20, 256, 32, 32)), ("15b x192 ResidualBlock", "ResidualBlock", "Normal", (15, 192, 32, 32))):
    stack = [base + ("-SE" if (i + 1) % 3 == 0 else "") for i in range(shape[0])]
    path = os.path.join(tempfile.gettempdir(), "fam_speed.bin"); synth.write_synth_net(path, shape, seed=3, stack=stack, policy_head=head)
    for prec in (0, 1):
        pipe = engine.B200ForwardPipe().initialize(path, 19, b, gpus=[0], precision=prec)
        planes = [pos[i % 64] for i in range(b)]
        pipe.batch_forward(0, planes, [19]*b, [0]*b)
        pipe.time_forward(0, 0, 5, flush_l2=True)
        ms, _, _ = pipe.time_forward(0, 0, 20, flush_l2=True)
        print("%-34s precision %d batch %d: %.3f ms, %.0f evals/s" % (name, prec, b, float(np.median(ms)), b / float(np.median(ms)) * 1e3), flush=True)
        pipe.destroy()
PY
