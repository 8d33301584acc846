"""Shape fuzz: small towers of many widths / head sizes / block families / board sizes through the C ABI against the
CPU oracle (1e-4 on the fp32-split rung), to exercise every N-tile width, K-block count and padding path of the
convolution kernel.  Run on the GPU box:  python tools/fuzz_parity.py"""
import itertools
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_py  # noqa: E402
from sayuri_b200 import engine, synth  # noqa: E402

oracle_py.build()
rng = np.random.default_rng(12345)
widths = [16, 32, 48, 64, 80, 96, 112, 128, 144, 160, 192, 224, 256]
heads = [(8, 8), (16, 16), (24, 24), (32, 32), (8, 24), (40, 24), (16, 48)]
families = ["ResidualBlock", "BottleneckBlock", "NestedBottleneckBlock", "MixerBlock"]
acts = ["mish", "relu", "swish"]
worst = 0.0
n_cfg = 0
fails = []
for C in widths:
    for fam in families:
        P, V = heads[int(rng.integers(len(heads)))]
        act = acts[int(rng.integers(len(acts)))]
        head = "RepLK" if rng.random() < 0.4 else "Normal"
        k = int(rng.choice([3, 5, 7, 9]))
        stack = [fam + ("-SE" if rng.random() < 0.5 else ""), "ResidualBlock" + ("-SE" if rng.random() < 0.5 else ""), fam]
        path = os.path.join(tempfile.gettempdir(), "fuzz.bin")
        try:
            synth.write_synth_net(path, (3, C, P, V), seed=int(rng.integers(1 << 30)), stack=stack, activation=act, policy_head=head, dw_kernel=max(k, 3))
        except Exception as ex:
            print("skip (writer)", C, fam, ex)
            continue
        tag = "C=%d %s P=%d V=%d %s head=%s k=%d" % (C, "/".join(stack), P, V, act, head, k)
        try:
            pipe = engine.B200ForwardPipe().initialize(path, 19, 40, gpus=[0])
        except RuntimeError as ex:
            print("REJECTED", tag, "->", str(ex)[:120])
            continue
        try:
            orc = oracle_py.Oracle(path)
            for batch in (1, 5, 40):
                sizes = [int(rng.choice([19, 19, 13, 9, 7, 2])) for _ in range(batch)]
                planes = [synth.synth_positions(1, bs, seed=int(rng.integers(1 << 30)))[0].ravel() for bs in sizes]
                offs = [int(rng.integers(5)) for _ in range(batch)]
                out = pipe.batch_forward(0, planes, sizes, offs)
                for i in sorted(set([0, batch - 1, int(rng.integers(batch))])):
                    ref = orc.forward(planes[i], sizes[i], offs[i])
                    s = sizes[i] ** 2
                    d = max(float(np.abs(out[i]["probabilities"][:s] - ref["prob"]).max()), float(np.abs(out[i]["ownership"][:s] - ref["own"]).max()),
                            abs(float(out[i]["pass_probability"]) - float(ref["misc"][0])), float(np.abs(np.asarray(out[i]["wdl"]) - ref["misc"][1:4]).max()))
                    worst = max(worst, d)
                    if not d < 1e-4:
                        fails.append((tag, batch, i, sizes[i], d))
            n_cfg += 1
            print("ok  %-110s worst so far %.2e" % (tag, worst), flush=True)
        finally:
            pipe.destroy()
print("configs %d, max |cuda - oracle| = %.3g, failures: %d" % (n_cfg, worst, len(fails)))
for f in fails[:20]:
    print("FAIL", f)
sys.exit(1 if fails else 0)
