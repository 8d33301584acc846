"""the sharded NN result cache (sayuri_b200/csrc/shim/utils/cache.h) that shadows the reference's
utils/cache.h in the front-end build.  Bit-exact observable behaviour (hits, misses, evictions, values, growth, clear,
move) against digests produced by the REFERENCE's header on the same operation sequences (tests/golden/
make_cache_golden.py), live against the reference build when oracle/_ref is present, and a concurrent run that must
never return a torn or foreign value."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "sayuri_b200", "csrc", "shim")
HARNESS_SRC = os.path.join(ROOT, "oracle", "cache_harness.cc")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "cache_harness_ref")


@pytest.fixture(scope="module")
def ours(tmp_path_factory):
    """oracle/cache_harness.cc compiled against OUR header only (no reference tree on the include path)."""
    out = str(tmp_path_factory.mktemp("cache") / "cache_harness_b200")
    subprocess.run(["g++", "-std=c++17", "-O2", "-pthread", "-I" + SHIM, HARNESS_SRC, "-o", out], check=True)
    return out


def _digest(binary, n_ops, seed, cap):
    return subprocess.run([binary, "parity", str(n_ops), str(seed), str(cap)], check=True, capture_output=True, text=True).stdout.strip()


def test_cache_matches_reference_digests(ours):
    with open(os.path.join(ROOT, "tests", "golden", "cache_digests.json")) as f:
        cases = json.load(f)
    assert len(cases) >= 10
    for c in cases:
        assert _digest(ours, c["n_ops"], c["seed"], c["capacity"]) == c["digest"], c


def test_cache_matches_reference_build_live_when_present(ours):
    if not os.path.exists(REF_BIN):
        pytest.skip("oracle/_ref not built")
    for seed in (3, 5, 8):
        for cap in (16, 777, 12345):
            assert _digest(ours, 50000, seed, cap) == _digest(REF_BIN, 50000, seed, cap), (seed, cap)


@pytest.mark.parametrize("threads,capacity,key_space,probes", [(16, 4096, 20000, 1), (64, 256, 1000, 8), (8, 8, 50, 2)])
def test_cache_concurrent_hits_are_whole_values_of_the_right_key(ours, threads, capacity, key_space, probes):
    r = subprocess.run([ours, "bench", str(threads), "0.7", str(capacity), str(key_space), str(probes)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    stats = json.loads(r.stdout)
    assert stats["torn"] == 0 and stats["hit_rate"] > 0.01 and stats["evals_per_s"] > 0
