"""Forward-throughput sweeps for the other BASELINE.json configurations (device-timed, inputs resident, L2 flushed):
config 5 (20bx256, 19x19, batch 1..2048), config 3's net at batch 512, config 4 (15bx192, mixed 9/13/19 batch).
Prints a markdown table; run on the GPU box:  python tools/sweep.py > gpurun_out/sweep.md"""
import argparse
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sayuri_b200 import engine, synth  # noqa: E402


def flops_per_eval(net, S=361):
    b, C, P, V = synth.NETS[net]
    return 2 * S * (9 * 43 * C + b * 2 * 9 * C * C + C * P + 5 * P + C * V + V)


def measure(net, batch, precision, sizes=None, iters=None):
    path = os.path.join(tempfile.gettempdir(), "sweep_%s.bin" % net)
    if not os.path.exists(path):
        synth.write_synth_net(path, net, seed=20260417)
    pipe = engine.B200ForwardPipe().initialize(path, 19, batch, gpus=[0], precision=precision)
    sizes = sizes or [19] * batch
    pos = {bs: synth.synth_positions(8, bs, seed=5).reshape(8, -1) for bs in set(sizes)}
    planes = [pos[bs][i % 8] for i, bs in enumerate(sizes)]
    pipe.batch_forward(0, planes, sizes, [0] * batch)
    iters = iters or max(5, min(60, int(6000 / max(batch, 16))))
    pipe.time_forward(0, 0, 3, flush_l2=True)
    ms, _, _ = pipe.time_forward(0, 0, iters, flush_l2=True)
    host_ev, _ = pipe.time_batch_forward_host(0, planes, sizes, [0] * batch, 0.6)   # blocking call, pageable host buffers
    pipe.destroy()
    med = float(np.median(ms))
    return batch / med * 1e3, med, host_ev


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    print("| config | net | batch | rung | ms / forward | evals/s (device) | algorithmic TFLOP/s | evals/s through sb_forward_batch (host buffers) |")
    print("|---|---|---|---|---|---|---|---|")
    batches = [1, 4, 16, 64, 256, 1024, 2048] if a.quick else [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048]
    for prec, name in ((engine.PRECISION_FP32_SPLIT, "fp32-split"), (engine.PRECISION_FP16, "fp16")):
        for b in batches:
            ev, ms, hev = measure("20bx256", b, prec)
            print("| 5 | 20bx256 | %d | %s | %.3f | %.0f | %.1f | %.0f |" % (b, name, ms, ev, ev * flops_per_eval("20bx256") / 1e12, hev), flush=True)
    for prec, name in ((engine.PRECISION_FP32_SPLIT, "fp32-split"), (engine.PRECISION_FP16, "fp16")):
        sizes = [(9, 13, 19)[i % 3] for i in range(256)]
        ev, ms, hev = measure("15bx192", 256, prec, sizes=sizes)
        fl = sum(flops_per_eval("15bx192", S=s * s) for s in sizes) / 256
        print("| 4 | 15bx192 | 256 (mixed 9/13/19) | %s | %.3f | %.0f | %.1f | %.0f |" % (name, ms, ev, ev * fl / 1e12, hev), flush=True)
        ev, ms, hev = measure("20bx256", 512, prec)
        print("| 3 | 20bx256 | 512 | %s | %.3f | %.0f | %.1f | %.0f |" % (name, ms, ev, ev * flops_per_eval("20bx256") / 1e12, hev), flush=True)
        ev, ms, hev = measure("10bx128", 256, prec)
        print("| 2 | 10bx128 | 256 | %s | %.3f | %.0f | %.1f | %.0f |" % (name, ms, ev, ev * flops_per_eval("10bx128") / 1e12, hev), flush=True)
        ev, ms, hev = measure("10bx128", 1024, prec)
        print("| 2 | 10bx128 | 1024 | %s | %.3f | %.0f | %.1f | %.0f |" % (name, ms, ev, ev * flops_per_eval("10bx128") / 1e12, hev), flush=True)
        ev, ms, hev = measure("6bx96", 256, prec, sizes=[9] * 256)
        print("| 1 | 6bx96 | 256 (9x9 on 19x19 canvas) | %s | %.3f | %.0f | %.1f | %.0f |" % (name, ms, ev, ev * flops_per_eval("6bx96", 81) / 1e12, hev), flush=True)


if __name__ == "__main__":
    main()
