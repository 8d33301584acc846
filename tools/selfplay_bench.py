"""Self-play games/hour of the UNMODIFIED reference self-play loop (src/selfplay/pipe.cc:235-296, engine.cc) — the second
half of BASELINE.json's metric.  The loop runs in ONE process over all listed GPUs (`-g 0 -g 1 ...`, as the reference
front-end drives several GPUs), over our pipe (oracle/_ref/sayuri_b200_frontend = reference front-end + the C++ shim +
libsayuri_b200.so) or over the reference's own Eigen CPU pipe (oracle/_ref/sayuri_eigen_v3) for the baseline beside it.
Every game is played to its end; games/hour = finished games / wall time of the loop (from "backend ready" to exit).

    python tools/selfplay_bench.py --net 10bx128 --board 19 --playouts 400 --parallel-games 128 --gpus 0
    python tools/selfplay_bench.py --preset config3 --gpus 0,1,2,3,4,5,6,7
    python tools/selfplay_bench.py --preset config4 --gpus 0,1,2,3
    python tools/selfplay_bench.py --binary sayuri_eigen_v3 --net 10bx128 --parallel-games 16 --games 16     # CPU baseline

Prints one JSON line.  Nothing here reads /root/reference: the binaries are built by oracle/Makefile and shipped."""
import argparse
import glob
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sayuri_b200 import synth  # noqa: E402

PRESETS = {
    # BASELINE.json configs[1] as self-play: 19x19, 10bx128, 400 visits
    "config2": dict(net="10bx128", queries=["bkp:19:7:1.0"], playouts=400, extra=[]),
    # configs[2]: 19x19, 20bx256, 800 visits, 512 parallel games over 8 GPUs
    "config3": dict(net="20bx256", queries=["bkp:19:7:1.0"], playouts=800, extra=[], parallel_games=512),
    # configs[3]: mixed 9/13/19 boards, 15bx192, Gumbel-MCTS on 4 GPUs; search settings of the reference's own
    # bash/configs/selfplay-config.txt (Gumbel with 150 playouts, fast searches of 50 with probability 0.75)
    "config4": dict(net="15bx192", queries=["bkp:9:7:0.334", "bkp:13:7:0.333", "bkp:19:7:0.333"], playouts=150,
                    extra=["--gumbel", "--gumbel-playouts-threshold", "32", "--gumbel-prom-visits", "1",
                           "--fastsearch-playouts", "50", "--fastsearch-playouts-prob", "0.75",
                           "--random-fastsearch-prob", "0.75", "--dirichlet-noise", "--early-symm-cache", "--first-pass-bonus"]),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--preset", choices=sorted(PRESETS))
    ap.add_argument("--net", default="10bx128")
    ap.add_argument("--board", type=int, default=19)
    ap.add_argument("--playouts", type=int, default=400)
    ap.add_argument("--parallel-games", type=int, default=0, help="default 128 per GPU (16 on the CPU pipe)")
    ap.add_argument("--games", type=int, default=0, help="default = parallel games (every thread plays one game)")
    ap.add_argument("--gpus", default="0")
    ap.add_argument("--batch-size", type=int, default=0, help="0 = the reference's default: parallel_games / (2 x GPUs)")
    ap.add_argument("--fp16", action="store_true", help="the front-end's own default precision (fp16 rung); default here is --no-fp16 = the parity rung")
    ap.add_argument("--binary", default="sayuri_b200_frontend")
    ap.add_argument("--cache-mib", type=int, default=2000)
    ap.add_argument("--timeout", type=float, default=3000)
    ap.add_argument("--window", type=float, default=0.0,
                    help="> 0: stop the loop after this many seconds and report the NN evaluation rate only (games do NOT finish: "
                         "no games/hour); for configurations whose games would take longer than the GPU budget allows")
    ap.add_argument("--label", default="")
    ap.add_argument("--extra", default="", help="further front-end flags, space separated")
    a = ap.parse_args()

    cfg = dict(net=a.net, queries=["bkp:%d:7:1.0" % a.board], playouts=a.playouts, extra=[])
    if a.preset:
        cfg.update(PRESETS[a.preset])
    exe = os.path.join(ROOT, "oracle", "_ref", a.binary)
    if not os.path.exists(exe):
        raise SystemExit("%s is not built (make -C oracle where /root/reference exists)" % exe)
    cpu_pipe = "b200" not in a.binary
    gpus = [] if cpu_pipe else [int(g) for g in a.gpus.split(",") if g != ""]
    pg = a.parallel_games or cfg.get("parallel_games") or (16 if cpu_pipe else 128 * max(1, len(gpus)))
    games = a.games or pg
    weights = os.path.join(tempfile.gettempdir(), "sp_%s.bin" % cfg["net"])
    if not os.path.exists(weights):
        synth.write_synth_net(weights + ".tmp", cfg["net"], seed=20260417)
        os.replace(weights + ".tmp", weights)
    out = tempfile.mkdtemp(prefix="sp_out_")
    cmd = [exe, "--mode", "selfplay", "-w", weights, "--parallel-games", str(pg), "--num-games", str(games),
           "-p", str(cfg["playouts"]), "--target-directory", out, "--cache-memory-mib", str(a.cache_mib)]
    for q in cfg["queries"]:
        cmd += ["--selfplay-query", q]
    if not cpu_pipe:
        cmd += [] if a.fp16 else ["--no-fp16"]   # fp16 is the front-end's default (config.cc:33); there is no --fp16 flag
        for g in gpus:
            cmd += ["-g", str(g)]
    if a.batch_size > 0:
        cmd += ["-b", str(a.batch_size)]
    cmd += cfg["extra"] + a.extra.split()

    env = dict(os.environ)
    stats_file = os.path.join(out, "engine_stats.txt")
    if a.window > 0:
        env["SAYURI_B200_STATS_FILE"] = stats_file
    t0 = time.perf_counter()
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    ready = [None]
    lines = []

    def pump():
        for line in proc.stdout:
            lines.append(line.rstrip())
            # the backend announces itself when the engine (weights on every GPU, slots, batcher) is up
            if ready[0] is None and ("replica(s)" in line or "Version:" in line or "BLAS" in line):
                ready[0] = time.perf_counter()

    th = threading.Thread(target=pump, daemon=True)
    th.start()
    try:
        proc.wait(timeout=a.window if a.window > 0 else a.timeout)
        timed_out = False
    except subprocess.TimeoutExpired:
        proc.kill()
        timed_out = True
    t1 = time.perf_counter()
    th.join(timeout=5)
    loop_s = t1 - (ready[0] or t0)
    n_sgf = sum(open(f, errors="replace").read().count("(;") for f in glob.glob(os.path.join(out, "sgf", "*")))
    q = 0
    for f in glob.glob(os.path.join(out, "net_queries", "*.txt")):
        rows = [x.split() for x in open(f).read().strip().splitlines() if x.strip()]
        if rows:
            q = max(q, int(rows[-1][-1]))
    finished = any("Totally played" in ln for ln in lines)
    moves = 0
    for f in glob.glob(os.path.join(out, "sgf", "*")):
        txt = open(f, errors="replace").read()
        moves += txt.count(";B[") + txt.count(";W[")
    res = {"what": "self-play through the unmodified reference loop", "label": a.label or (a.preset or "custom"),
           "binary": a.binary, "pipe": "reference Eigen CPU" if cpu_pipe else "sayuri_b200 (%s rung)" % ("fp16" if a.fp16 else "fp32-split"),
           "net": cfg["net"], "queries": cfg["queries"], "playouts": cfg["playouts"], "parallel_games": pg, "games_requested": games,
           "gpus": gpus, "host_cores": os.cpu_count(), "finished_all_games": bool(finished and not timed_out and n_sgf >= games),
           "games_finished": n_sgf, "moves_played": moves, "loop_seconds": round(loop_s, 2),
           "startup_seconds": round((ready[0] or t0) - t0, 2), "nn_evals": q,
           "games_per_hour": round(n_sgf * 3600.0 / loop_s, 1) if n_sgf else 0.0,
           "nn_evals_per_sec": round(q / loop_s, 1), "moves_per_sec": round(moves / loop_s, 1),
           "extra_flags": cfg["extra"] + a.extra.split(), "timed_out": timed_out, "returncode": proc.returncode}
    if a.window > 0:
        # rate over the second half of the window from the engine's own counters (first half = opening moves, cache warm-up)
        rows = [ln.split() for ln in open(stats_file)] if os.path.exists(stats_file) else []
        rows = [(float(r[0]), int(r[1]), int(r[2])) for r in rows if len(r) == 3]
        if len(rows) >= 4:
            mid, last = rows[len(rows) // 2], rows[-1]
            res["window"] = {"seconds": a.window, "nn_evals_per_sec_second_half": round((last[2] - mid[2]) / (last[0] - mid[0]), 1),
                             "mean_batch": round((last[2] - mid[2]) / max(1, last[1] - mid[1]), 1), "nn_evals_total": last[2],
                             "note": "window run: games were cut, games/hour is NOT measured"}
        res["games_per_hour"] = None
        if "window" not in res:
            res["log_tail"] = lines[-12:]
    elif not finished:
        res["log_tail"] = lines[-8:]
    shutil.rmtree(out, ignore_errors=True)
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
