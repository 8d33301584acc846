"""GPU bring-up probe: runs the engine on a ladder of synthetic nets / modes, each case in its own
subprocess (a trapped kernel poisons the CUDA context), and compares trunk + outputs with the CPU
oracle.  Usage on the GPU box:  python tools/gpu_probe.py [--cases a,b,...] > gpurun_out/probe.log
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# name -> (blocks, C, P, V, stack or None)
SHAPES = {
    "in32": (0, 32, 8, 8, []),
    "b1c64": (1, 64, 8, 8, ["ResidualBlock"]),
    "b1c128": (1, 128, 24, 24, ["ResidualBlock"]),
    "b2c128se": (2, 128, 24, 24, ["ResidualBlock", "ResidualBlock-SE"]),
    "b2c96se": (2, 96, 24, 24, ["ResidualBlock-SE", "ResidualBlock"]),
    "b2c192se": (2, 192, 32, 32, ["ResidualBlock", "ResidualBlock-SE"]),
    "b2c256se": (2, 256, 32, 32, ["ResidualBlock", "ResidualBlock-SE"]),
    "10bx128": (10, 128, 24, 24, None),
    "15bx192": (15, 192, 32, 32, None),
    "20bx256": (20, 256, 32, 32, None),
    "golden": None,
}
MODES = {"simt": (2, 0), "split": (0, 0), "fp16": (1, 0), "split2": (0, 2), "fp16_2": (1, 2)}


def run_case(shape, mode, n, seed):
    from oracle.oracle_py import Oracle
    from sayuri_b200 import engine, synth
    prec, bo = MODES[mode]
    if shape == "golden":
        path = os.path.join(ROOT, "tests", "golden", "ref_3bx32.bin.txt")
    else:
        blocks, C, P, V, stack = SHAPES[shape]
        path = os.path.join(tempfile.gettempdir(), "probe_%s.bin" % shape)
        synth.write_synth_net(path, (blocks, C, P, V), seed=5, stack=stack)
    sizes = [(19, 13, 9)[i % 3] if n > 1 else 19 for i in range(n)]
    if n >= 4:
        sizes[0] = 19
    planes = [synth.synth_positions(1, bs, seed=seed + i)[0].ravel() for i, bs in enumerate(sizes)]
    offsets = [i % 5 for i in range(n)]
    pipe = engine.B200ForwardPipe().initialize(path, 19, max(n, 4), gpus=[0], precision=prec)
    if bo == 2:
    out = pipe.batch_forward(0, planes, sizes, offsets)
    orc = Oracle(path)
    res = {"shape": shape, "mode": mode, "n": n, "trunk": 0.0, "prob": 0.0, "own": 0.0, "misc": 0.0, "scale": 0.0}
    worst = None
    for i, bs in enumerate(sizes):
        s = bs * bs
        ref = orc.forward_trace(planes[i], bs, offsets[i])
        trunk = pipe.debug_read_trunk(0, 0, i, bs)
        dt = float(np.abs(trunk - ref["trunk"]).max())
        if dt > res["trunk"]:
            res["trunk"] = dt
            worst = (i, bs, trunk, ref["trunk"])
        res["scale"] = max(res["scale"], float(np.abs(ref["trunk"]).max()))
        res["prob"] = max(res["prob"], float(np.abs(out[i]["probabilities"][:s] - ref["prob"]).max()))
        res["own"] = max(res["own"], float(np.abs(out[i]["ownership"][:s] - ref["own"]).max()))
        got_misc = np.array([out[i]["pass_probability"], *out[i]["wdl"], out[i]["stm_winrate"], out[i]["final_score"],
                             out[i]["q_error"], out[i]["score_error"]])
        res["misc"] = max(res["misc"], float(np.abs(got_misc - ref["misc"]).max()))
        tail = float(np.abs(out[i]["probabilities"][s:]).max()) if s < 361 else 0.0
        res["tail_nonzero"] = max(res.get("tail_nonzero", 0.0), tail)
    if worst is not None and res["trunk"] > 1e-3:
        i, bs, got, ref = worst
        err = np.abs(got - ref)
        bad = err > 1e-3
        res["diag"] = {
            "sample": i, "bs": bs, "frac_bad": float(bad.mean()),
            "bad_channels": int(bad.any(axis=1).sum()), "bad_pixels": int(bad.any(axis=0).sum()),
            "first_bad_pixels": np.nonzero(bad.any(axis=0))[0][:24].tolist(),
            "first_bad_channels": np.nonzero(bad.any(axis=1))[0][:24].tolist(),
            "got0": got[:2, :6].round(4).tolist(), "ref0": ref[:2, :6].round(4).tolist(),
            "nan": bool(np.isnan(got).any()),
        }
    print("RESULT " + json.dumps(res))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", default=None, help="internal: shape:mode:n")
    ap.add_argument("--shapes", default="in32,b1c64,b1c128,golden")
    ap.add_argument("--modes", default="simt,split,fp16")
    ap.add_argument("--n", default="5,16")
    args = ap.parse_args()
    if args.case:
        shape, mode, n = args.case.split(":")
        run_case(shape, mode, int(n), seed=100)
        return
    for shape in args.shapes.split(","):
        for mode in args.modes.split(","):
            for n in args.n.split(","):
                cmd = [sys.executable, os.path.abspath(__file__), "--case", "%s:%s:%s" % (shape, mode, n)]
                try:
                    p = subprocess.run(cmd, capture_output=True, text=True, timeout=180)
                    lines = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
                    if lines:
                        print(lines[-1][7:], flush=True)
                    else:
                        tail = (p.stderr or p.stdout).strip().splitlines()[-3:]
                        print(json.dumps({"shape": shape, "mode": mode, "n": int(n), "rc": p.returncode, "error": tail}), flush=True)
                except subprocess.TimeoutExpired:
                    print(json.dumps({"shape": shape, "mode": mode, "n": int(n), "error": "TIMEOUT"}), flush=True)


if __name__ == "__main__":
    main()
