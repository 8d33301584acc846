"""Generates tests/golden/cache_digests.json from the REFERENCE's cache (oracle/_ref/cache_harness_ref = oracle/cache_harness.cc
compiled against /root/reference/src/utils/cache.h by oracle/Makefile).  Run in the container that has /root/reference."""
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "..", "..", "oracle", "_ref", "cache_harness_ref")
CASES = [(n_ops, seed, cap) for cap in (8, 64, 1000, 4096, 50001) for n_ops, seed in ((20000, 1), (200000, 20260417))]


def main():
    out = []
    for n_ops, seed, cap in CASES:
        digest = subprocess.run([REF, "parity", str(n_ops), str(seed), str(cap)], check=True, capture_output=True, text=True).stdout.strip()
        out.append({"n_ops": n_ops, "seed": seed, "capacity": cap, "digest": digest})
    with open(os.path.join(HERE, "cache_digests.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(len(out), "digests written")


if __name__ == "__main__":
    main()
