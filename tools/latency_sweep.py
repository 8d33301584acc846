"""Device time of one forward at small batches (the GTP / analysis case), with an option on and off.
   python tools/latency_sweep.py --option layer_overlap --nets 10bx128,20bx256 --batches 1,4,16,64"""
import argparse
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sayuri_b200 import engine, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nets", default="10bx128,20bx256")
ap.add_argument("--batches", default="1,4,16,64")
ap.add_argument("--option", default="layer_overlap")
ap.add_argument("--values", default="0,1")
ap.add_argument("--iters", type=int, default=40)
a = ap.parse_args()
values = [int(v) for v in a.values.split(",")]
print("| net | rung | batch | " + " | ".join("%s=%d, us" % (a.option, v) for v in values) + " |")
print("|---|---|---|" + "---|" * len(values))
for net in a.nets.split(","):
    path = os.path.join(tempfile.gettempdir(), "lat_%s.bin" % net)
    synth.write_synth_net(path, net, seed=1)
    for prec, name in ((0, "split"), (1, "fp16")):
        for batch in [int(b) for b in a.batches.split(",")]:
            pipe = engine.B200ForwardPipe().initialize(path, 19, batch, gpus=[0], precision=prec)
            x = synth.synth_positions(batch, 19, seed=3).reshape(batch, -1)
            cells = []
            for v in values:
                pipe.set_option(a.option, v)
                pipe.batch_forward(0, list(x), [19] * batch, [0] * batch)
                ms, _, _ = pipe.time_forward(0, 0, a.iters, flush_l2=False)
                cells.append("%.1f" % (1e3 * float(np.median(ms))))
            pipe.destroy()
            print("| %s | %s | %d | %s |" % (net, name, batch, " | ".join(cells)))
