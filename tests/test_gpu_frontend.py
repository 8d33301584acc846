"""North-star observables THROUGH THE C++ SHIM (sayuri_b200/csrc/shim/b200_forward_pipe.cc = class CudaForwardPipe over the
C ABI), driven by the UNMODIFIED reference front-end (GTP + MCTS) — SURVEY.md §8 rows a3 (plugin interface) and a24
(callers: identical root visit counts under a fixed seed).

oracle/_ref/sayuri_b200_det  = reference front-end + our pipe   (fp32-split rung, --no-fp16)
oracle/_ref/sayuri_eigen_det = reference front-end + its own Eigen CPU pipe (the oracle of BASELINE.json's north star)
Both differ from the stock reference only in utils/random.cc (oracle/det_random.cc: seed from $SAYURI_SEED instead of
the thread id; the reference has no seed flag).  Built by oracle/Makefile from /root/reference, shipped to the GPU box
as binaries; nothing here reads /root/reference at run time."""
import concurrent.futures as cf
import os
import sys
import tempfile

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "sayuri_eigen_det")
OUR_BIN = os.path.join(ROOT, "oracle", "_ref", "sayuri_b200_det")


def _gtp(board, moves):
    return "boardsize %d\nclear_board\n" % board + "".join("genmove %s\n" % ("b" if i % 2 == 0 else "w") for i in range(moves)) + "quit\n"


def _compare(weights, board, playouts, moves, seeds, our_extra=(), our_env=None):
    """Root child visit vectors of every search, reference CPU pipe vs our pipe, per seed (searches after a divergent
    move are not comparable and would be dropped — none may occur)."""
    import visit_parity
    if not (os.path.exists(REF_BIN) and os.path.exists(OUR_BIN)):
        pytest.fail("oracle/_ref/sayuri_{eigen,b200}_det are not shipped: run `make -C oracle` where /root/reference exists")
    gtp = _gtp(board, moves)
    saved = dict(os.environ)
    if our_env:
        os.environ.update(our_env)   # visit_parity.run copies os.environ
    try:
        with cf.ThreadPoolExecutor(max_workers=min(len(seeds), max(2, (os.cpu_count() or 4) - 2))) as pool:
            ref_f = {s: pool.submit(visit_parity.run, REF_BIN, weights, gtp, playouts, s, []) for s in seeds}
            ours = {}
            for s in seeds:   # one engine at a time on the GPU
                ours[s] = visit_parity.run(OUR_BIN, weights, gtp, playouts, s, ["--no-fp16", "-g", "0", *our_extra])
            refs = {s: f.result() for s, f in ref_f.items()}
    finally:
        os.environ.clear()
        os.environ.update(saved)
    total, report = 0, []
    for s in seeds:
        (rs, rm, rout), (os_, om, oout) = refs[s], ours[s]
        assert "sayuri_b200" in oout, "our pipe did not announce itself:\n" + oout[-1500:]
        assert len(rs) == moves and len(os_) == moves, (s, len(rs), len(os_), oout[-1500:])
        for i, (x, y) in enumerate(zip(rs, os_)):
            diff = {k: (x.get(k, 0), y.get(k, 0)) for k in set(x) | set(y) if x.get(k, 0) != y.get(k, 0)}
            assert not diff, "seed %d search %d: root visit counts differ (reference, ours): %r" % (s, i, diff)
            assert sum(x.values()) > 0
            total += 1
        assert rm == om, (s, rm, om)
        report.append((s, rm))
    return total, report


def test_identical_root_visit_counts_19x19_400_playouts_through_the_shim():
    """>= 30 root searches on 19x19 at 400 playouts (BASELINE.json metric's visit count): every child's visit count and
    every chosen move equal the reference Eigen pipe's.  6bx96 net (BASELINE config 1's net) so that the single-threaded
    CPU arm finishes in about a minute per seed."""
    from sayuri_b200 import synth
    w = os.path.join(tempfile.gettempdir(), "sb_vp_6bx96.bin")
    synth.write_synth_net(w, "6bx96", seed=11)
    total, report = _compare(w, 19, 400, 6, [1, 2, 3, 4, 5, 6])
    assert total >= 30
    with open(os.path.join(tempfile.gettempdir(), "sb_visit_parity.log"), "a") as f:
        f.write("19x19 6bx96 -p 400: %d root searches, all visit vectors and moves identical: %r\n" % (total, report))


def test_mixed_board_9x9_game_on_a_19x19_engine_through_batchforward():
    """A 9x9 game evaluated on a 19x19 NN canvas (--fixed-nn-boardsize 19) through the reference's own batcher in front
    of CudaForwardPipe::BatchForward (SAYURI_B200_REF_BATCHER=1): the shim un-lays the host canvas the reference built
    (batch_forward_pipe.cc:15-33), the engine places / masks / crops on the device, and the search still reproduces the
    Eigen pipe's visit counts (which runs the net at the native 9x9)."""
    from sayuri_b200 import synth
    w = os.path.join(tempfile.gettempdir(), "sb_vp_10bx128.bin")
    synth.write_synth_net(w, "10bx128", seed=11)
    total, report = _compare(w, 9, 200, 4, [7, 8, 9], our_extra=["--fixed-nn-boardsize", "19"],
                             our_env={"SAYURI_B200_REF_BATCHER": "1"})
    assert total == 12
    with open(os.path.join(tempfile.gettempdir(), "sb_visit_parity.log"), "a") as f:
        f.write("9x9 on a 19x19 canvas, reference batcher + BatchForward, 10bx128 -p 200: %d root searches identical: %r\n" % (total, report))


def test_engine_batcher_path_of_the_shim_matches_on_13x13():
    """Same check through CudaForwardPipe::Forward -> sb_eval (the default path of the shim), 13x13 on its own canvas."""
    from sayuri_b200 import synth
    w = os.path.join(tempfile.gettempdir(), "sb_vp_6bx96.bin")
    synth.write_synth_net(w, "6bx96", seed=11)
    total, _ = _compare(w, 13, 300, 4, [21, 22])
    assert total == 8
