#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "== wide N (fp16, 256-wide layers) on/off"
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/wide_n.log
import os, sys, tempfile
sys.path.insert(0, '.')
import numpy as np
from sayuri_b200 import engine, synth
from oracle import oracle_py
oracle_py.build()
pos = synth.synth_positions(64, 19, seed=5).reshape(64, -1)
path = os.path.join(tempfile.gettempdir(), "wn_20bx256.bin"); synth.write_synth_net(path, "20bx256", seed=20260417)
orc = oracle_py.Oracle(path)
ref = orc.forward(pos[3], 19, 0)
for b in (8, 256, 1024):
    pipe = engine.B200ForwardPipe().initialize(path, 19, b, gpus=[0], precision=1)
    planes = [pos[i % 64] for i in range(b)]
    outs = {}
    for wn in (0, 1, 0, 1):
        pipe.set_option("wide_n", wn)
        pipe.reload(path)
        out = pipe.batch_forward(0, planes, [19]*b, [0]*b)
        outs[wn] = out
        pipe.time_forward(0, 0, 3, flush_l2=True)
        ms, cms, cn = pipe.time_forward(0, 0, 12, flush_l2=True, profile_conv=True)
        err = float(np.abs(out[3]["probabilities"] - ref["prob"]).max())
        print("20bx256 fp16 batch %d wide_n %d: %.3f ms, %.0f evals/s (convs %.3f ms) | max |policy - oracle| %.2e" % (b, wn, float(np.median(ms)), b / float(np.median(ms)) * 1e3, cms, err), flush=True)
    print("   bit-identical wide vs narrow:", all(np.array_equal(outs[0][f], outs[1][f]) for f in ("probabilities", "ownership", "wdl")))
    pipe.destroy()
PY
echo "== bench fp16 20bx256"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision fp16 --net 20bx256 --eval-threads 0 2>&1 | tail -1 | cut -c1-300
echo "== bench"; timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300 | tee gpurun_out/bench.log
