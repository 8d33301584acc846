"""Host-side ceiling of sb_eval: a tiny net (GPU time negligible), many native threads.  evals/s here / host cores = what
one core sustains through pack + claim + sleep + wake + copy-out."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sayuri_b200 import engine, synth
path = os.path.join(tempfile.gettempdir(), "tiny.bin")
synth.write_synth_net(path, (1, 16, 8, 8), seed=1, stack=["ResidualBlock"])
pos = synth.synth_positions(256, 19, seed=5).reshape(256, -1)
for batch, wait in ((256, 200), (64, 50), (1024, 200)):
    pipe = engine.B200ForwardPipe().initialize(path, 19, batch, gpus=[0], precision=1)
    pipe.batcher_config(batch, wait)
    for t in (16, 64, 256, 1024, 4096):
        before = pipe.batcher_stats()
        ev = pipe.eval_throughput(pos, 19, t, 2.0)
        st = pipe.batcher_stats()
        nb = st["batches"] - before["batches"]
        print("tiny net, batch<=%d wait %dus, %4d threads on %d cores: %8.0f evals/s (%.1f us of core time per eval), mean batch %.1f" % (
            batch, wait, t, os.cpu_count(), ev, 1e6 * os.cpu_count() / ev, (st["positions"] - before["positions"]) / max(nb, 1)), flush=True)
    pipe.destroy()
