"""MCTS visit-count parity under a fixed seed: the UNMODIFIED reference front-end (GTP + MCTS) is run twice on the
same positions with the same SAYURI_SEED — once over the reference Eigen CPU pipe (oracle/_ref/sayuri_eigen_det),
once over our pipe (oracle/_ref/sayuri_b200_det, fp32-split rung) — and the root child visit vectors are compared.
Both binaries differ from the stock reference only in utils/random.cc (oracle/det_random.cc: seed from the
environment instead of the thread id).  Prints one JSON summary line."""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sayuri_b200 import synth  # noqa: E402


def run(binary, weights, gtp, playouts, seed, extra):
    env = dict(os.environ, SAYURI_SEED=str(seed))
    cmd = [binary, "-w", weights, "-t", "1", "-b", "1", "-p", str(playouts), "-a", "--no-winograd"] + extra
    p = subprocess.run(cmd, input=gtp, capture_output=True, text=True, timeout=3000, env=env)
    out = p.stdout + p.stderr
    searches, cur = [], None
    for line in out.splitlines():
        if re.match(r"\s*move\s+visits", line):
            cur = {}
            searches.append(cur)
            continue
        if cur is not None:
            m = re.match(r"\s*([A-T]\d+|pass)\s+(\d+)\s", line, re.I)
            if m:
                cur[m.group(1)] = int(m.group(2))
            elif line.strip().startswith("* Tree"):
                cur = None
    moves = re.findall(r"^= ([A-T]\d+|pass|resign)\s*$", out, re.M | re.I)
    return searches, moves, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--net", default="6bx96")
    ap.add_argument("--board", type=int, default=9)
    ap.add_argument("--playouts", type=int, default=400)
    ap.add_argument("--moves", type=int, default=6)
    ap.add_argument("--seeds", default="1,2,3")
    ap.add_argument("--weights", default=None, help="an existing weight file instead of a synthetic --net")
    ap.add_argument("--host-override", action="store_true",
                    help="compare sayuri_eigen_det with sayuri_eigen_det_fast (same Eigen pipe, link-time "
                         "Board::ComputePassAliveArea override) instead of our pipe: runs without a GPU")
    a = ap.parse_args()
    if a.weights:
        w = a.weights
        a.net = os.path.basename(w)
    else:
        w = os.path.join(tempfile.gettempdir(), "vp_%s.bin" % a.net)
        synth.write_synth_net(w, a.net, seed=11)
    gtp = "boardsize %d\nclear_board\n" % a.board + "".join("genmove %s\n" % ("b" if i % 2 == 0 else "w") for i in range(a.moves)) + "quit\n"
    ref_bin = os.path.join(ROOT, "oracle", "_ref", "sayuri_eigen_det")
    our_bin = os.path.join(ROOT, "oracle", "_ref", "sayuri_eigen_det_fast" if a.host_override else "sayuri_b200_det")
    total = same = 0
    first_div = []
    l1 = []
    for seed in [int(s) for s in a.seeds.split(",")]:
        rs, rm, _ = run(ref_bin, w, gtp, a.playouts, seed, [])
        os_, om, oout = run(our_bin, w, gtp, a.playouts, seed, [] if a.host_override else ["--no-fp16", "-g", "0"])
        if not os_:
            print(oout[-2000:])
            raise SystemExit("our front-end produced no search output")
        diverged = None
        for i, (x, y) in enumerate(zip(rs, os_)):
            if diverged is not None:
                break          # after a different move the positions differ: not comparable
            total += 1
            keys = set(x) | set(y)
            d = sum(abs(x.get(k, 0) - y.get(k, 0)) for k in keys)
            l1.append(d)
            if d == 0:
                same += 1
            if i < len(rm) and i < len(om) and rm[i] != om[i]:
                diverged = i
        first_div.append(diverged)
    print(json.dumps({"net": a.net, "board": a.board, "playouts": a.playouts, "searches_compared": total,
                      "identical_visit_vectors": same, "l1_visit_diff_per_search": l1, "first_divergent_move_per_seed": first_div}))


if __name__ == "__main__":
    main()
