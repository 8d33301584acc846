"""Throughput of the reference-facing single-position call (sb_eval = NetworkForwardPipe::Forward behind the shim)
driven by T native host threads on synthetic positions, fp32 planes in pageable memory: packing on the calling
thread, H2D, forward, D2H and wake-up are all inside the timed region.
    python tools/eval_bench.py --net 10bx128 --threads 64,256,512 --batch 256"""
import argparse
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sayuri_b200 import engine, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--net", default="10bx128")
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--threads", default="32,128,512")
ap.add_argument("--seconds", type=float, default=3.0)
ap.add_argument("--wait-us", type=int, default=200)
ap.add_argument("--precision", type=int, default=0)
ap.add_argument("--gpus", default="", help="replicas of ONE engine in this process, e.g. 0,1,2,3 (default: all devices)")
ap.add_argument("--async-depth", type=int, default=0, help="> 0: feeder threads with this many tickets in flight each (sb_eval_submit / sb_eval_wait)")
a = ap.parse_args()
path = os.path.join(tempfile.gettempdir(), "evalbench_%s.bin" % a.net)
synth.write_synth_net(path, a.net, seed=20260417)
pos = synth.synth_positions(512, 19, seed=5).reshape(512, -1)
gpus = [int(g) for g in a.gpus.split(",") if g != ""] or None
pipe = engine.B200ForwardPipe().initialize(path, 19, a.batch, gpus=gpus, precision=a.precision)
print("replicas %d, weights: %r" % (pipe.get_num_workers(), pipe.weights_stats()), flush=True)
pipe.batcher_config(a.batch, a.wait_us)
for t in [int(v) for v in a.threads.split(",")]:
    before = pipe.batcher_stats()
    ev = pipe.eval_throughput_async(pos, 19, t, a.async_depth, a.seconds) if a.async_depth > 0 else pipe.eval_throughput(pos, 19, t, a.seconds)
    st = pipe.batcher_stats()
    nb = st["batches"] - before["batches"]
    print("net %s precision %d batch<=%d wait %dus %s %4d: %9.0f evals/s  (mean batch %.1f, %d full / %d timer closes, %d GPU workers)" % (
        a.net, a.precision, a.batch, a.wait_us, ("feeders x %d tickets" % a.async_depth) if a.async_depth > 0 else "threads", t, ev, (st["positions"] - before["positions"]) / max(nb, 1),
        st["full"] - before["full"], st["timer"] - before["timer"], st["workers"]), flush=True)
pipe.destroy()
