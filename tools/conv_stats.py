"""Print the conv3x3_tc in-kernel cycle counters (last conv launch of a forward) for a net / batch."""
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
ap.add_argument("--dbg", type=int, default=0)
ap.add_argument("--option", action="append", default=[])
ap.add_argument("--tail-split", type=int, default=1)
ap.add_argument("--launch", type=int, default=2, help="conv launch index within the forward (2 = second tower conv)")
a = ap.parse_args()
path = os.path.join(tempfile.gettempdir(), "stats_%s.bin" % a.net)
synth.write_synth_net(path, a.net, seed=1)
pipe = engine.B200ForwardPipe().initialize(path, 19, a.batch, gpus=[0], precision=a.precision)
x = synth.synth_positions(min(a.batch, 32), 19, seed=3).reshape(-1, engine.PLANE_FLOATS)
planes = [x[i % x.shape[0]] for i in range(a.batch)]
pipe.set_option("tail_split", a.tail_split)
for kv in a.option:
    pipe.set_option(kv.split("=")[0], int(kv.split("=")[1]))
pipe.batch_forward(0, planes, [19] * a.batch, [0] * a.batch)
pipe.set_option("stats", 1)
pipe.set_option("stats_launch", a.launch)
pipe.set_option("conv_dbg", a.dbg)
pipe.batch_forward(0, planes, [19] * a.batch, [0] * a.batch)
st = pipe.conv_stats(0, 0).astype(np.float64)
st = st[st[:, 6] > 0]
names = ["mma_total", "wait_tmem_empty", "wait_slab", "wait_b", "epi_wait_full", "epi_total", "items", "epi_drain"]
ms, conv_ms, conv_n = pipe.time_forward(0, 0, 10, flush_l2=True, profile_conv=True)
print("net %s batch %d precision %d dbg %d tail_split %d launch %d: %d clusters; forward %.3f ms, convs %.3f ms over %d launches" % (
    a.net, a.batch, a.precision, a.dbg, a.tail_split, a.launch, st.shape[0], float(np.median(ms)), conv_ms, conv_n))
for i, n in enumerate(names):
    col = st[:, i]
    print("  %-16s mean %10.0f  min %10.0f  max %10.0f" % (n, col.mean(), col.min(), col.max()))
for n_it in sorted(set(st[:, 6].astype(int))):
    sel = st[st[:, 6] == n_it]
    print("  pairs with %d units: %3d   mma_total mean %8.0f   epi_total mean %8.0f" % (n_it, sel.shape[0], sel[:, 0].mean(), sel[:, 5].mean()))
busy = st[:, 0] - st[:, 1] - st[:, 2] - st[:, 3]
print("  mma issue+exec (total - waits) per item: mean %.0f cycles" % (busy / np.maximum(st[:, 6], 1)).mean())
