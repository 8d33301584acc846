"""Where does the error of the fp32-split rung come from?  Trunk outputs of the tensor-core path (tcgen05, fp32 accumulators
in TMEM), of the SIMT fp32 cross-check kernel on the SAME hi/lo buffers, and of the CPU oracle (plain fp32), for towers of
growing width and depth.  Prints, per net: max / rms error of TC and SIMT against the oracle, and the signed statistics of
(TC - SIMT) that separate a multiplicative shrink (truncating accumulation) from an additive offset and from noise."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle_py import Oracle  # noqa: E402
from sayuri_b200 import engine, synth  # noqa: E402


def trunks(path, planes, prec, chunk=1, comp=0):
    pipe = engine.B200ForwardPipe().initialize(path, 19, 4, gpus=[0], precision=prec)
    try:
        pipe.set_option("chunk_accumulate", chunk)
        pipe.set_option("acc_comp_ppb", comp)
        pipe.batch_forward(0, planes, [19] * len(planes), [0] * len(planes))
        return [pipe.debug_read_trunk(0, 0, i, 19) for i in range(len(planes))]
    finally:
        pipe.destroy()


SETTINGS = [(0, 0), (0, 8), (0, 12), (1, 0), (1, 12)]   # (chunk_accumulate, acc_comp_ppb)


def main():
    planes = [synth.synth_positions(1, 19, seed=4000 + 17 * i)[0].ravel() for i in range(2)]
    for C, blocks in ((128, 10), (96, 6), (128, 20), (192, 15)):
        path = os.path.join(tempfile.gettempdir(), "sb_prec_%d_%d.bin" % (C, blocks))
        synth.write_synth_net(path, (blocks, C, 32, 32), seed=20260417, stack=["ResidualBlock"] * blocks)
        orc = Oracle(path)
        ref = [orc.forward_trace(p, 19, 0)["trunk"] for p in planes]
        si = trunks(path, planes, engine.PRECISION_SIMT_DEBUG)
        r = np.concatenate([x.ravel() for x in ref]).astype(np.float64)
        s = np.concatenate([x.ravel() for x in si]).astype(np.float64)
        rms = np.sqrt(np.mean(r * r))
        print("C=%d blocks=%d  |trunk| rms %.3g max %.3g | SIMT-oracle max %.3g rms %.3g" % (
            C, blocks, rms, np.abs(r).max(), np.abs(s - r).max(), np.sqrt(np.mean((s - r) ** 2))), flush=True)
        for ct, comp in SETTINGS:
            tc = trunks(path, planes, engine.PRECISION_FP32_SPLIT, ct, comp)
            t = np.concatenate([x.ravel() for x in tc]).astype(np.float64)
            d = t - s
            a, b = np.polyfit(s, d, 1)   # least-squares fit d = a * s + b
            print("    chunk_accumulate %d comp %3d ppb: TC-oracle max %.3g rms %.3g | TC-SIMT slope %.3g offset %.3g resid-rms %.3g"
                  % (ct, comp, np.abs(t - r).max(), np.sqrt(np.mean((t - r) ** 2)), a, b, np.sqrt(np.mean((d - a * s - b) ** 2))), flush=True)


if __name__ == "__main__":
    main()
