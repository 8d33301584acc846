#!/bin/bash
# Is self-play bound by host CPU or by the number of leaves in flight?  Same loop, same net, 9x9, 100 visits, with 64 and
# 128 parallel (single-threaded) games.  tools/selfplay_concurrency.sh [timeout_s]
mkdir -p gpurun_out
python -c "
import sys; sys.path.insert(0,'.')
from sayuri_b200 import synth
synth.write_synth_net('/tmp/fe_10bx128.bin', '10bx128', seed=20260417)"
for PG in 64 128; do
  rm -rf /tmp/spc && mkdir -p /tmp/spc
  S=$(date +%s.%N)
  timeout ${1:-20} oracle/_ref/sayuri_b200_frontend --mode selfplay -w /tmp/fe_10bx128.bin --no-fp16 -g 0 --parallel-games $PG --num-games $PG -p 100 \
     --selfplay-query bkp:9:7:1.0 --target-directory /tmp/spc --cache-memory-mib 1000 > /tmp/spc.log 2>&1
  E=$(date +%s.%N)
  python - <<PY
import glob
t = $E - $S
q = 0
for f in glob.glob('/tmp/spc/net_queries/*.txt'):
    lines = [x.split() for x in open(f).read().strip().splitlines() if x.strip()]
    if lines: q = max(q, int(lines[-1][-1]))
n = sum(open(f).read().count('(;') for f in glob.glob('/tmp/spc/sgf/*'))
print("9x9 10bx128 -p 100, $PG parallel games: %d games in %.1f s -> %.0f games/hour, %.0f NN evals/s" % (n, t, n * 3600 / t, q / t))
PY
done | tee gpurun_out/selfplay_concurrency.log
