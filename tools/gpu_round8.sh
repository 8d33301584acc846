#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-400 | tee gpurun_out/bench.log
echo "== tower throughput by block family (10 blocks x 128, SE every 3rd, batch 256)"
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/family_speed.log
import os, sys, tempfile
sys.path.insert(0, '.')
import numpy as np
from sayuri_b200 import engine, synth
pos = synth.synth_positions(64, 19, seed=5).reshape(64, -1)
b = 256
for name, base, head in (("ResidualBlock", "ResidualBlock", "Normal"), ("BottleneckBlock", "BottleneckBlock", "Normal"),
                         ("NestedBottleneckBlock", "NestedBottleneckBlock", "Normal"), ("MixerBlock + RepLK head", "MixerBlock", "RepLK")):
    stack = [base + ("-SE" if (i + 1) % 3 == 0 else "") for i in range(10)]
    path = os.path.join(tempfile.gettempdir(), "fam_speed.bin"); synth.write_synth_net(path, (10, 128, 24, 24), seed=3, stack=stack, policy_head=head)
    for prec in (0, 1):
        pipe = engine.B200ForwardPipe().initialize(path, 19, b, gpus=[0], precision=prec)
        planes = [pos[i % 64] for i in range(b)]
        pipe.batch_forward(0, planes, [19]*b, [0]*b)
        pipe.time_forward(0, 0, 5, flush_l2=True)
        ms, _, _ = pipe.time_forward(0, 0, 20, flush_l2=True)
        print("10b x128 %-26s precision %d batch %d: %.3f ms, %.0f evals/s" % (name, prec, b, float(np.median(ms)), b / float(np.median(ms)) * 1e3), flush=True)
        pipe.destroy()
PY
