"""BASELINE config 1 as plumbing: the reference's own `--mode selfplay` loop finishes 9x9 games and
writes the SGF record, the training-data chunks and the NN-query log — with the unmodified reference
(sayuri_eigen_v3) and with all host-side replacements of DESIGN.md §5b (sayuri_eigen_fast: sharded NN cache,
pass-alive and reach area, encoder planes, data-writer thread), over the reference's Eigen CPU pipe so that it runs
without a GPU.
The same loop over our pipe is what tools/selfplay_host.sh times on the B200."""
import glob
import os
import resource
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


def test_data_writer_replacement_does_not_spin():
    """The reference's writer thread polls with sched_yield for the whole run (pipe.cc:191-192); the replacement sleeps
    between polls: same records, a fraction of the system time."""
    ref, fast = (os.path.join(ROOT, "oracle", "_ref", b) for b in ("sayuri_eigen_v3", "sayuri_eigen_fast"))
    if not (os.path.exists(ref) and os.path.exists(fast)):
        pytest.skip("oracle/_ref not built")
    import tempfile
    from sayuri_b200 import synth
    sys_time = {}
    with tempfile.TemporaryDirectory() as d:
        weights = os.path.join(d, "tiny.bin")
        synth.write_synth_net(weights, (1, 16, 8, 8), seed=3)
        for exe in (ref, fast):
            out = os.path.join(d, os.path.basename(exe))
            os.mkdir(out)
            before = resource.getrusage(resource.RUSAGE_CHILDREN)
            r = subprocess.run([exe, "--mode", "selfplay", "-w", weights, "--parallel-games", "1", "--num-games", "2", "-p", "100",
                                "--selfplay-query", "bkp:9:7:1.0", "--target-directory", out, "--cache-memory-mib", "50"],
                               capture_output=True, text=True, timeout=900)
            after = resource.getrusage(resource.RUSAGE_CHILDREN)
            assert r.returncode == 0 and "Totally played 2 games" in r.stdout + r.stderr
            assert len(glob.glob(os.path.join(out, "tdata", "*", "*"))) == 2
            sys_time[exe] = after.ru_stime - before.ru_stime
    assert sys_time[ref] > 0.3, sys_time          # the spin is there in the reference ...
    assert sys_time[fast] < 0.25 * sys_time[ref], sys_time   # ... and gone in the replacement
