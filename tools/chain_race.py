"""Race hunt for the chained / overlapped convolution launches: BASELINE config 2 at full batch, the same positions in two
orders, outputs compared bit for bit, repeated; prints where they differ.  python tools/chain_race.py --option conv_chain=1"""
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
ap.add_argument("--precision", type=int, default=0)
ap.add_argument("--repeats", type=int, default=6)
ap.add_argument("--option", action="append", default=[])
a = ap.parse_args()
path = os.path.join(tempfile.gettempdir(), "race_%s.bin" % a.net)
synth.write_synth_net(path, a.net, seed=20260417)
n = a.batch
x = synth.synth_positions(n, 19, seed=20260419).reshape(n, -1)
pipe = engine.B200ForwardPipe().initialize(path, 19, n, gpus=[0], precision=a.precision)
base_opts = dict(kv.split("=") for kv in a.option)
for k in base_opts:
    pipe.set_option(k, 0)   # reference run: everything named on the command line off
pipe.set_option("layer_overlap", 0)
pipe.set_option("conv_chain", 0)
ref = pipe.batch_forward(0, list(x), [19] * n, [0] * n)
for k, v in base_opts.items():
    pipe.set_option(k, int(v))
rng = np.random.default_rng(0)
bad = 0
for rep in range(a.repeats):
    perm = rng.permutation(n) if rep else np.arange(n)
    out = pipe.batch_forward(0, list(x[perm]), [19] * n, [0] * n)
    for f in ("probabilities", "ownership"):
        d = out[f] != ref[f][perm]
        if d.any():
            bad += 1
            pos = np.unique(np.nonzero(d)[0])
            print("rep %d %s: %d elements differ in %d positions (batch slots %s...), max |diff| %.3g" % (
                rep, f, int(d.sum()), pos.size, pos[:8].tolist(), float(np.abs(out[f] - ref[f][perm])[d].max())))
print("%s batch %d precision %d options %s: %d of %d comparisons differ" % (a.net, n, a.precision, a.option, bad, 2 * a.repeats))
pipe.destroy()
