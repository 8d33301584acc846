#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
echo "== small-batch N split on/off"
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/small_batch.log
import os, sys, tempfile
sys.path.insert(0, '.')
import numpy as np
from sayuri_b200 import engine, synth
for net in ("20bx256", "10bx128", "6bx96"):
    path = os.path.join(tempfile.gettempdir(), "sbs_%s.bin" % net); synth.write_synth_net(path, net, seed=20260417)
    pos = synth.synth_positions(64, 19, seed=5).reshape(64, -1)
    for prec in (0, 1):
        for b in (1, 2, 4, 8, 16, 32):
            pipe = engine.B200ForwardPipe().initialize(path, 19, b, gpus=[0], precision=prec)
            planes = [pos[i % 64] for i in range(b)]
            ref = None
            res = []
            for sbs in (0, 1):
                pipe.set_option("small_batch_split", sbs)
                out = pipe.batch_forward(0, planes, [19]*b, [0]*b)
                if ref is None: ref = out
                same = all(np.array_equal(out[f], ref[f]) for f in ("probabilities", "ownership", "wdl"))
                pipe.time_forward(0, 0, 5, flush_l2=True)
                ms, _, _ = pipe.time_forward(0, 0, 30, flush_l2=True)
                res.append((float(np.median(ms)), same))
            print("%s precision %d batch %2d: split off %.4f ms (%.0f evals/s) | on %.4f ms (%.0f evals/s) bit-identical %s" % (
                net, prec, b, res[0][0], b / res[0][0] * 1e3, res[1][0], b / res[1][0] * 1e3, res[1][1]), flush=True)
            pipe.destroy()
PY
echo "== chain_forwards on/off: pipelined submit/wait and sb_eval"
timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/chain.log
import os, sys, tempfile, time
sys.path.insert(0, '.')
import numpy as np
from sayuri_b200 import engine, synth
path = os.path.join(tempfile.gettempdir(), "ch_10bx128.bin"); synth.write_synth_net(path, "10bx128", seed=20260417)
B = 256
pos = synth.synth_positions(64, 19, seed=5).reshape(64, -1)
for prec in (0, 1):
    pipe = engine.B200ForwardPipe().initialize(path, 19, B, gpus=[0], precision=prec)
    pinned = engine.PinnedArray((2, B, engine.PLANE_FLOATS))
    for k in range(2):
        for i in range(B): pinned.array[k, i] = pos[(i + 7 * k) % 64]
    outs = [np.zeros(B, dtype=engine.OUTPUT_DTYPE) for _ in range(2)]
    sizes, offs = [19] * B, [0] * B
    def loop(steps):
        for s in range(steps):
            slot = s & 1
            if s >= 2: pipe.wait(0, slot, outs[slot])
            pipe.submit(0, slot, pinned.array[slot], sizes, offs)
        for s in range(max(steps - 2, 0), steps): pipe.wait(0, s & 1, outs[s & 1])
    for chain in (0, 1, 0, 1):
        pipe.set_option("chain_forwards", chain)
        loop(6)
        t0 = time.perf_counter(); loop(60); t = time.perf_counter() - t0
        ev = pipe.eval_throughput(pos, 19, 512, 1.5)
        print("precision %d chain_forwards %d: pipelined submit/wait %.0f evals/s, sb_eval(512 threads) %.0f evals/s" % (prec, chain, B * 60 / t, ev), flush=True)
    pinned.free(); pipe.destroy()
PY
echo "== bench"; timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1500 | tee gpurun_out/bench.log
echo "== bench fp16"; timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --precision fp16 2>&1 | tail -1 | cut -c1-1500 | tee gpurun_out/bench_fp16.log
