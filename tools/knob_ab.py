"""A/B of one scheduling knob (sb_set_option key) on device-timed forwards, interleaved on/off so that clock and power
drift hits both arms alike.  Run on the GPU box:  python tools/knob_ab.py --knob pdl_aux"""
import argparse
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sayuri_b200 import engine, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--knob", default="pdl_aux")
    ap.add_argument("--rounds", type=int, default=6)
    ap.add_argument("--iters", type=int, default=40)
    a = ap.parse_args()
    cases = [("10bx128", 256), ("10bx128", 32), ("20bx256", 256), ("6bx96", 256)]
    print("| net | batch | rung | ms off | ms on | on/off |")
    print("|---|---|---|---|---|---|")
    for net, batch in cases:
        path = os.path.join(tempfile.gettempdir(), "ab_%s.bin" % net)
        if not os.path.exists(path):
            synth.write_synth_net(path, net, seed=20260417)
        for prec, name in ((engine.PRECISION_FP32_SPLIT, "fp32-split"), (engine.PRECISION_FP16, "fp16")):
            pipe = engine.B200ForwardPipe().initialize(path, 19, batch, gpus=[0], precision=prec)
            pos = synth.synth_positions(8, 19, seed=5).reshape(8, -1)
            planes = [pos[i % 8] for i in range(batch)]
            pipe.batch_forward(0, planes, [19] * batch, [0] * batch)
            pipe.time_forward(0, 0, 5, flush_l2=True)
            t = {0: [], 1: []}
            for r in range(a.rounds):
                for v in (0, 1):
                    pipe.set_option(a.knob, v)
                    pipe.time_forward(0, 0, 2, flush_l2=True)
                    ms, _, _ = pipe.time_forward(0, 0, a.iters, flush_l2=True)
                    t[v].append(float(np.median(ms)))
            off, on = float(np.median(t[0])), float(np.median(t[1]))
            print("| %s | %d | %s | %.4f | %.4f | %.4f |" % (net, batch, name, off, on, on / off), flush=True)
            pipe.destroy()


if __name__ == "__main__":
    main()
