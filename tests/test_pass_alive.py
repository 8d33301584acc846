"""the pass-alive / pass-dead area (sayuri_b200/csrc/host_go/pass_alive.h) that replaces
Board::ComputePassAliveArea at link time in the front-end build.  Bit-exact against the reference's answers: committed
fixtures (no reference needed), and live — function against function, and front-end callers against front-end callers
through the actual link-time override — when oracle/_ref is present."""
import gzip
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


def test_pass_alive_replays_reference_fixtures(tmp_path):
    exe, raw = str(tmp_path / "replay"), str(tmp_path / "cases.bin")
    subprocess.run(["g++", "-std=c++17", "-O2", os.path.join(ROOT, "tests", "pass_alive_replay.cc"), "-o", exe], check=True)
    with gzip.open(os.path.join(ROOT, "tests", "golden", "pass_alive_cases.bin.gz"), "rb") as g, open(raw, "wb") as f:
        f.write(g.read())
    r = subprocess.run([exe, raw], capture_output=True, text=True)
    stats = json.loads(r.stdout)
    assert r.returncode == 0 and stats["mismatches"] == 0, stats
    assert stats["records"] > 2000 and stats["answers"] == 8 * stats["records"] and stats["marked_points"] > 100000
    assert stats["reach_answers"] == stats["records"]


def _record(rows, answers, reach):
    """One fixture record (format of tests/golden/make_pass_alive_golden.py) from ASCII rows: X black, O white, . empty."""
    n = len(rows)
    code = {"X": 0, "O": 1, ".": 2}
    rec = bytes([n, 0]) + bytes(code[c] for row in rows for c in row)
    for a in answers:
        rec += bytes(1 if c == "#" else 0 for row in a for c in row)
    return rec + bytes(code[c] for row in reach for c in row)


def test_pass_alive_hand_worked_position(tmp_path):
    """A position small enough to work by hand.  Black: one string with three one-point eyes on the edge => alive
    (two healthy vital regions suffice); the open area below is nobody's.  White has no stones."""
    rows = [".X.X.",
            "XXXXX",
            ".....",
            ".....",
            "....."]
    stones = ["-#-#-", "#####", "-----", "-----", "-----"]
    with_eyes = ["#####", "#####", "-----", "-----", "-----"]
    nothing = ["-----"] * 5
    # colour-major, flags = vitals | dead << 1.  Black: 0 = living stones; 1 = + their vital eyes; 2 = the eyes are
    # reported as pass-dead regions of the opponent (no eye of his fits there); 3 = both.  White: nothing at all.
    answers = [stones, with_eyes, with_eyes, with_eyes, nothing, nothing, nothing, nothing]
    reach = ["XXXXX"] * 5          # only black stones on the board: every point is reached by black alone
    exe, raw = str(tmp_path / "replay"), str(tmp_path / "one.bin")
    subprocess.run(["g++", "-std=c++17", "-O2", os.path.join(ROOT, "tests", "pass_alive_replay.cc"), "-o", exe], check=True)
    with open(raw, "wb") as f:
        f.write(_record(rows, answers, reach))
    r = subprocess.run([exe, raw], capture_output=True, text=True)
    assert r.returncode == 0 and json.loads(r.stdout)["mismatches"] == 0, r.stdout


def test_pass_alive_hand_worked_dead_group(tmp_path):
    """Black lives with two edge eyes and walls in a white group that has a single eye: the white group and its eye are
    black's pass-dead region; white itself has only one vital region, so nothing of white is alive; the open right side
    is left alone (room for three and more white eyes)."""
    rows = [".X.X..",
            "XXXX..",
            "OOOX..",
            ".OOX..",
            "OOOX..",
            "XXXX.."]
    stones = ["-#-#--", "####--", "---#--", "---#--", "---#--", "####--"]
    alive = ["####--", "####--", "---#--", "---#--", "---#--", "####--"]
    with_dead = ["####--", "####--", "####--", "####--", "####--", "####--"]
    nothing = ["------"] * 6
    answers = [stones, alive, with_dead, with_dead, nothing, nothing, nothing, nothing]
    reach = ["XXXXXX", "XXXXXX", "OOOXXX", "OOOXXX", "OOOXXX", "XXXXXX"]
    exe, raw = str(tmp_path / "replay"), str(tmp_path / "one.bin")
    subprocess.run(["g++", "-std=c++17", "-O2", os.path.join(ROOT, "tests", "pass_alive_replay.cc"), "-o", exe], check=True)
    with open(raw, "wb") as f:
        f.write(_record(rows, answers, reach))
    r = subprocess.run([exe, raw], capture_output=True, text=True)
    assert r.returncode == 0 and json.loads(r.stdout)["mismatches"] == 0, r.stdout


def test_pass_alive_matches_reference_function_live_when_present():
    exe = os.path.join(REF, "pass_alive_harness")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref not built")
    r = subprocess.run([exe, "check", "63", "5"], capture_output=True, text=True)
    stats = json.loads(r.stdout)
    assert r.returncode == 0 and stats["mismatches"] == 0 and stats["calls_with_pass_dead_points"] > 100, (stats, r.stderr[-2000:])


def test_link_time_override_gives_the_same_safe_and_score_areas_when_present():
    plain, fast = os.path.join(REF, "pass_alive_harness"), os.path.join(REF, "pass_alive_harness_fast")
    if not (os.path.exists(plain) and os.path.exists(fast)):
        pytest.skip("oracle/_ref not built")
    # the override is really linked: the reference's symbol is weak in that binary's board object, ours is the strong one
    for seed in ("1", "9"):
        a = subprocess.run([plain, "digest", "42", seed], check=True, capture_output=True, text=True).stdout.strip()
        b = subprocess.run([fast, "digest", "42", seed], check=True, capture_output=True, text=True).stdout.strip()
        assert len(a) == 16 and a == b
    t = json.loads(subprocess.run([fast, "time", "4", "3", "19"], check=True, capture_output=True, text=True).stdout)
    # in the override build "the reference's" member function IS ours: both columns time the same code
    assert t["reference_ns_per_call"] < 4 * t["ours_ns_per_call"]


def test_whole_encoder_is_bit_identical_under_the_link_time_overrides_when_present():
    """Encoder::GetPlanes (encoder.cc:31-50), all eight symmetries, on positions of random games: the build with every
    override linked (pass-alive, reach area, symmetry gather, stone and last-move planes) hashes the same float bits."""
    plain, fast = os.path.join(REF, "pass_alive_harness"), os.path.join(REF, "pass_alive_harness_fast")
    if not (os.path.exists(plain) and os.path.exists(fast)):
        pytest.skip("oracle/_ref not built")
    a = json.loads(subprocess.run([plain, "encoder", "21", "4"], check=True, capture_output=True, text=True).stdout)
    b = json.loads(subprocess.run([fast, "encoder", "21", "4"], check=True, capture_output=True, text=True).stdout)
    assert a["encoded_positions"] == b["encoded_positions"] > 3000 and a["digest"] == b["digest"]


def test_ladder_map_replays_reference_fixtures(tmp_path):
    """sb_go::LadderMap (host_go/ladder.h, the link-time replacement of Board::GetLadderMap) on 933 positions of seeded
    fight-heavy random games, board sizes 5..19, against the REFERENCE's answers stored with them.  No reference needed."""
    exe, raw = str(tmp_path / "replay"), str(tmp_path / "cases.bin")
    subprocess.run(["g++", "-std=c++17", "-O2", os.path.join(ROOT, "tests", "ladder_replay.cc"), "-o", exe], check=True)
    with gzip.open(os.path.join(ROOT, "tests", "golden", "ladder_cases.bin.gz"), "rb") as g, open(raw, "wb") as f:
        f.write(g.read())
    r = subprocess.run([exe, raw], capture_output=True, text=True)
    stats = json.loads(r.stdout)
    assert r.returncode == 0 and stats["mismatches"] == 0, stats
    assert stats["records"] > 900 and stats["records_19x19"] > 300 and stats["marked_points"] > 4000


def test_ladder_map_matches_reference_function_live_when_present():
    """Function against function in one process (the unmodified Board::GetLadderMap vs sb_go::LadderMap on a copy of the
    board's string arrays), then the front-end's caller through the actual link-time override: same digest of all maps."""
    plain, fast = os.path.join(REF, "pass_alive_harness"), os.path.join(REF, "pass_alive_harness_fast")
    if not (os.path.exists(plain) and os.path.exists(fast)):
        pytest.skip("oracle/_ref not built")
    a = json.loads(subprocess.run([plain, "ladder", "150", "3"], capture_output=True, text=True).stdout)
    assert a["mismatches"] == 0 and a["positions"] > 30000 and a["positions_with_a_ladder"] > 10000, a
    b = json.loads(subprocess.run([fast, "ladder", "150", "3"], capture_output=True, text=True).stdout)
    assert b["mismatches"] == 0 and b["digest"] == a["digest"] and b["positions"] == a["positions"]
    # in the override build the member function is ours: no slower than the header called directly (same code)
    assert b["ns_per_map_member"] < 2.5 * b["ns_per_map_header"]
