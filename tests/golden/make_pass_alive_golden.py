"""Generates tests/golden/pass_alive_cases.bin.gz: positions of seeded random games (board sizes 2..19) with the answers of
the REFERENCE's Board::ComputePassAliveArea (/root/reference/src/game/board.cc:1720) for both colours and the four
(mark_vitals, mark_pass_dead) combinations, written by oracle/_ref/pass_alive_harness (oracle/pass_alive_harness.cc,
linked against the unmodified reference objects).  Run in the container that has /root/reference.
Record: u8 board_size, u8 0, n*n stone bytes (0 black, 1 white, 2 empty), then 8 answers of n*n bytes
(colour-major, flags = vitals | dead << 1), then the answer of Board::ComputeReachArea (board.cc:1547), n*n bytes."""
import gzip
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(HERE, "..", "..", "oracle", "_ref", "pass_alive_harness")


def main():
    with tempfile.TemporaryDirectory() as d:
        raw = os.path.join(d, "cases.bin")
        print(subprocess.run([BIN, "dump", "42", "20260417", raw], check=True, capture_output=True, text=True).stdout.strip())
        with open(raw, "rb") as f, gzip.GzipFile(os.path.join(HERE, "pass_alive_cases.bin.gz"), "wb", compresslevel=9, mtime=0) as g:
            g.write(f.read())


if __name__ == "__main__":
    main()
