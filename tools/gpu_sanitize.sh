#!/bin/bash
# compute-sanitizer memcheck over one small mixed-size forward of every block family (tiny nets, batch 6)
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import os, sys, tempfile
sys.path.insert(0, '.')
import numpy as np
from sayuri_b200 import engine, synth
sizes = [19, 13, 9, 19, 19, 13]
planes = [synth.synth_positions(1, bs, seed=40 + i)[0].ravel() for i, bs in enumerate(sizes)]
for tag, path in (("residual", "tests/golden/ref_3bx32.bin.txt"), ("bottleneck", "tests/golden/ref_btl_5bx32.bin.txt"), ("mixer+replk", "tests/golden/ref_mix_4bx32.bin.txt")):
    for prec in (0, 1):
        pipe = engine.B200ForwardPipe().initialize(path, 19, 8, gpus=[0], precision=prec)
        out = pipe.batch_forward(0, planes, sizes, [0, 1, 2, 3, 4, 0])
        one = pipe.eval(planes[0], 19, 0)
        assert np.isfinite(out["probabilities"]).all() and np.array_equal(one["probabilities"], out[0]["probabilities"])
        pipe.destroy()
        print("ok", tag, prec, flush=True)
PY
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python /tmp/san.py 2>&1 | tail -25 | tee gpurun_out/sanitizer_memcheck.log
echo "exit code: ${PIPESTATUS[0]}" | tee -a gpurun_out/sanitizer_memcheck.log
