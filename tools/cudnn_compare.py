"""BASELINE.json config 5: forward throughput sweep, batch 1..2048, 19x19, ours vs the UNMODIFIED reference cuDNN
backend (oracle/_ref/sayuri_cudnn_bench = CudaForwardPipe::BatchForward driven directly), same synthetic weights and
positions, host buffers in / host results out on both sides (the reference's pinned staging + H2D/D2H, ours through
sb_forward_batch).  Also compares the two pipes' raw outputs on the same positions.  Run on the GPU box:
    python tools/cudnn_compare.py --net 20bx256 > gpurun_out/cudnn_compare.md"""
import argparse
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sayuri_b200 import engine, synth  # noqa: E402

# SAYURI_CUDNN_BENCH=sayuri_cudnn_bench_ptx90 selects the as-shipped build (SIMT kernels JIT-compiled from compute_90 PTX)
REF = os.path.join(ROOT, "oracle", "_ref", os.environ.get("SAYURI_CUDNN_BENCH", "sayuri_cudnn_bench"))


def ours(path, pos, batch, precision, seconds):
    pipe = engine.B200ForwardPipe().initialize(path, 19, batch, gpus=[0], precision=precision)
    planes = [pos[i % len(pos)] for i in range(batch)]
    sizes = [19] * batch
    offs = [0] * batch
    out = None
    for _ in range(3):
        out = pipe.batch_forward(0, planes, sizes, offs)
    # host-timed, blocking C-ABI call with host buffers (same shape of measurement as the reference harness)
    ev, ms = pipe.time_batch_forward_host(0, planes, sizes, offs, seconds)
    dms, _, _ = pipe.time_forward(0, 0, max(5, min(40, int(4000 / max(batch, 16)))), flush_l2=True)
    pipe.destroy()
    return ev, ms, float(np.median(dms)), out


def reference(path, planes_path, n_pos, fp16, seconds, batches, out_path):
    cmd = [REF, path, planes_path, str(n_pos), str(int(fp16)), str(seconds), out_path] + [str(b) for b in batches]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1200)
    res = {}
    for m in re.finditer(r"batch=(\d+) iters=\d+ ms_per_forward=([\d.]+) evals_per_s=([\d.]+)", r.stdout):
        res[int(m.group(1))] = (float(m.group(3)), float(m.group(2)))
    if not res:
        sys.stderr.write(r.stdout[-2000:] + r.stderr[-2000:])
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--net", default="20bx256")
    ap.add_argument("--seconds", type=float, default=1.5)
    ap.add_argument("--batches", default="1,2,4,8,16,32,64,128,256,512,1024,2048")
    a = ap.parse_args()
    batches = [int(x) for x in a.batches.split(",")]
    tmp = tempfile.gettempdir()
    path = os.path.join(tmp, "cmp_%s.bin" % a.net)
    synth.write_synth_net(path, a.net, seed=20260417)
    n_pos = 64
    pos = synth.synth_positions(n_pos, 19, seed=5).reshape(n_pos, -1).astype(np.float32)
    planes_path = os.path.join(tmp, "cmp_planes.bin")
    pos.tofile(planes_path)

    ref = {}
    for fp16 in (0, 1):
        t0 = time.time()
        ref[fp16] = reference(path, planes_path, n_pos, fp16, a.seconds, batches, os.path.join(tmp, "cmp_ref%d.bin" % fp16))
        sys.stderr.write("reference fp16=%d done in %.0f s\n" % (fp16, time.time() - t0))

    print("# %s, 19x19: forward throughput through host buffers, ours vs the reference cuDNN backend (same B200)" % a.net)
    print()
    print("| batch | ref cuDNN fp32 evals/s | ref cuDNN fp16 evals/s | ours fp32-split evals/s (device-only) | ours fp16 evals/s (device-only) | ours-split / ref-fp32 | ours-fp16 / ref-fp16 |")
    print("|---|---|---|---|---|---|---|")
    last = {}
    for b in batches:
        row = {}
        for prec in (engine.PRECISION_FP32_SPLIT, engine.PRECISION_FP16):
            ev, ms, dms, out = ours(path, pos, b, prec, a.seconds)
            row[prec] = (ev, b / dms * 1e3)
            last[prec] = out
        r32 = ref[0].get(b, (float("nan"),))[0]
        r16 = ref[1].get(b, (float("nan"),))[0]
        s, h = row[engine.PRECISION_FP32_SPLIT], row[engine.PRECISION_FP16]
        print("| %d | %.0f | %.0f | %.0f (%.0f) | %.0f (%.0f) | %.2f | %.2f |" % (b, r32, r16, s[0], s[1], h[0], h[1], s[0] / r32, h[0] / r16), flush=True)

    # output agreement on the last batch size (first n positions): the reference GPU pipe vs ours
    print()
    nb = min(n_pos, batches[-1])
    for fp16, prec, name in ((0, engine.PRECISION_FP32_SPLIT, "ours fp32-split vs ref cuDNN fp32"), (1, engine.PRECISION_FP16, "ours fp16 vs ref cuDNN fp16")):
        f = os.path.join(tmp, "cmp_ref%d.bin" % fp16)
        if not os.path.exists(f):
            continue
        r = np.fromfile(f, dtype=np.float32).reshape(-1, 361 * 2 + 8)[:nb]
        o = last[prec]
        dp = max(float(np.abs(o[i]["probabilities"][:361] - r[i, :361]).max()) for i in range(nb))
        do = max(float(np.abs(o[i]["ownership"][:361] - r[i, 361:722]).max()) for i in range(nb))
        dw = max(float(np.abs(np.asarray(o[i]["wdl"]) - r[i, 723:726]).max()) for i in range(nb))
        print("* %s: max |d policy logit| %.3g, |d ownership| %.3g, |d wdl| %.3g over %d positions" % (name, dp, do, dw, nb))


if __name__ == "__main__":
    main()
