"""TEST INFRASTRUCTURE — evaluates positions with the UNMODIFIED compiled reference (oracle/_ref, Eigen
BlasForwardPipe, im2col path) in a process of its own: the reference keeps its options in a process-global map, so
one process can hold ONE net (oracle_py.Reference).  Falls back to the plain-C oracle when oracle/_ref was not shipped.

    python tests/ref_forward_worker.py <weights> <in.npz> <out.npz>
in.npz : planes_<i> float32 [43*bs*bs], sizes int32 [n], offsets int32 [n]
out.npz: out_<i> float32 [2*bs*bs + 8] = prob | own | misc(8), kind = "reference" | "port"
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    weights, src, dst = sys.argv[1:4]
    from oracle import oracle_py
    d = np.load(src)
    sizes, offsets = d["sizes"], d["offsets"]
    if oracle_py.Reference.available():
        net, kind = oracle_py.Reference(weights, winograd=False), "reference"
    else:
        net, kind = oracle_py.Oracle(weights), "port"
    out = {"kind": np.array(kind)}
    for i, (bs, off) in enumerate(zip(sizes, offsets)):
        r = net.forward(d["planes_%d" % i], int(bs), int(off))
        out["out_%d" % i] = np.concatenate([r["prob"], r["own"], r["misc"]]).astype(np.float32)
    np.savez(dst, **out)


if __name__ == "__main__":
    main()
