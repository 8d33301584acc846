"""BASELINE config 1 as plumbing: the reference's own `--mode selfplay` loop finishes a 9x9 game and
writes an SGF record, a training-data chunk and the NN-query log — here with the two host-side replacements of
DESIGN.md §5b linked in (sharded NN cache is header-only and not in this Eigen build; the link-time
Board::ComputePassAliveArea is), over the reference's Eigen CPU pipe so that it runs without a GPU.  The same loop
over our pipe is what tools/selfplay_host.sh times on the B200."""
import glob
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("binary", ["sayuri_eigen_fast", "sayuri_eigen_v3"])
def test_selfplay_loop_finishes_a_game_and_writes_its_records(tmp_path, binary):
    exe = os.path.join(ROOT, "oracle", "_ref", binary)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref not built")
    from sayuri_b200 import synth
    weights = str(tmp_path / "tiny.bin")
    synth.write_synth_net(weights, (1, 16, 8, 8), seed=3)
    out = tmp_path / "out"
    out.mkdir()
    r = subprocess.run([exe, "--mode", "selfplay", "-w", weights, "--parallel-games", "2", "--num-games", "2", "-p", "16",
                        "--selfplay-query", "bkp:9:7:1.0", "--target-directory", str(out), "--cache-memory-mib", "50"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    assert "Totally played 2 games" in r.stdout + r.stderr
    sgf = glob.glob(str(out / "sgf" / "*.sgf"))
    assert len(sgf) == 1 and open(sgf[0]).read().count("(;") == 2
    assert sum(os.path.getsize(f) for f in glob.glob(str(out / "tdata" / "*" / "*")) + glob.glob(str(out / "vdata" / "*" / "*"))) > 1000
    queries = [line.split() for f in glob.glob(str(out / "net_queries" / "*.txt")) for line in open(f) if line.strip()]
    assert len(queries) == 2 and int(queries[-1][-1]) > 100
