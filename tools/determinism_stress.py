"""The same small batch evaluated over and over: every output must equal the first one bit for bit (race hunt for the
kernels that overlap layers at small batches).  python tools/determinism_stress.py --repeats 1500"""
import argparse
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sayuri_b200 import engine, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--net", default="10bx128")
ap.add_argument("--batches", default="1,2,8")
ap.add_argument("--precision", type=int, default=0)
ap.add_argument("--repeats", type=int, default=1500)
ap.add_argument("--configs", default="default;conv_chain=0;layer_overlap=0;conv_chain=0,layer_overlap=0;conv_chain=1,layer_overlap=2")
a = ap.parse_args()
path = os.path.join(tempfile.gettempdir(), "stress_%s.bin" % a.net)
synth.write_synth_net(path, a.net, seed=20260417)
FIELDS = ("probabilities", "ownership", "pass_probability", "wdl")
for n in [int(b) for b in a.batches.split(",")]:
    x = synth.synth_positions(n, 19, seed=77).reshape(n, -1)
    first = None
    for cfg in a.configs.split(";"):
        pipe = engine.B200ForwardPipe().initialize(path, 19, n, gpus=[0], precision=a.precision)
        if cfg != "default":
            for kv in cfg.split(","):
                pipe.set_option(kv.split("=")[0], int(kv.split("=")[1]))
        ref = pipe.batch_forward(0, list(x), [19] * n, [0] * n)
        if first is None:
            first = ref
        bad = 0
        worst = 0.0
        for rep in range(a.repeats):
            out = pipe.batch_forward(0, list(x), [19] * n, [0] * n)
            d = any(not np.array_equal(out[f], ref[f]) for f in FIELDS)
            if d:
                bad += 1
                worst = max(worst, max(float(np.abs(np.asarray(out[f], dtype=np.float64) - np.asarray(ref[f], dtype=np.float64)).max()) for f in FIELDS))
        same_as_first = all(np.array_equal(ref[f], first[f]) for f in FIELDS)
        print("batch %d  %-40s %d of %d forwards differ from the first (max |diff| %.3g); first forward equals the default config's: %s" % (
            n, cfg, bad, a.repeats, worst, same_as_first), flush=True)
        pipe.destroy()
