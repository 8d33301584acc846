#!/bin/bash
# First GPU call of the next round (≈ 14 GPU-minutes on one B200): the measurements this round could not finish.
#   1. parity gate + the bench line (both rungs)
#   2. 19x19, 10bx128, 400 visits self-play through the unmodified loop with ALL host-side replacements of DESIGN.md 5b,
#      at 64 / 128 / 256 parallel games (the 64-game run of round 1 was cut at 185 s: 31.4 k NN evals/s, no games/h)
#   3. the same at 128 parallel games with sayuri_b200_frontend_refcache = our pipe under the reference's host code
#      with NO host-side replacement (A/B arm; at 64 parallel games that arm gave 618 games/h in round 1)
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 300 python -m pytest tests -x -q -m gpu 2>&1 | tail -2 | tee gpurun_out/next_pytest_gpu.log
echo "== bench (fp32-split, fp16)"
timeout 300 python bench.py --steps 50 --warmup 5 2>&1 | tail -1 > gpurun_out/next_bench_split.json; cut -c1-200 gpurun_out/next_bench_split.json
timeout 200 python bench.py --steps 50 --warmup 5 --precision fp16 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/next_bench_fp16.json; cut -c1-200 gpurun_out/next_bench_fp16.json
echo "== self-play 19x19"
: > gpurun_out/next_selfplay19.log
for PG in 64 128 256; do tools/selfplay_host.sh 19 $PG 330 | tee -a gpurun_out/next_selfplay19.log; done
tools/selfplay_host.sh 19 128 330 oracle/_ref/sayuri_b200_frontend_refcache | tee -a gpurun_out/next_selfplay19.log
