#!/bin/bash
# A/B of the sharded NN cache (sayuri_b200/csrc/shim/utils/cache.h) against the reference's utils/cache.h: the same
# UNMODIFIED self-play loop over our pipe, 9x9 (short games, 8 symmetry probes per leaf during the first 9 moves),
# 10bx128, 400 visits; plus the cache micro-benchmark on this box's cores.  Usage: tools/selfplay_cache_ab.sh [games] [parallel]
mkdir -p gpurun_out
python -c "
import sys; sys.path.insert(0,'.')
from sayuri_b200 import synth
synth.write_synth_net('/tmp/fe_10bx128.bin', '10bx128', seed=20260417)"
NG=${1:-128}
PG=${2:-64}
run() {
  rm -rf /tmp/sp9 && mkdir -p /tmp/sp9
  S=$(date +%s.%N)
  timeout 200 $2 --mode selfplay -w /tmp/fe_10bx128.bin --no-fp16 -g 0 --parallel-games $PG --num-games $NG -p 400 \
     --selfplay-query bkp:9:7:1.0 --target-directory /tmp/sp9 --cache-memory-mib 2000 2>&1 | tail -1
  E=$(date +%s.%N)
  python - <<PY
import glob, os
t = $E - $S
q = 0
for f in glob.glob('/tmp/sp9/net_queries/*.txt'):
    lines = [x.split() for x in open(f).read().strip().splitlines() if x.strip()]
    if lines: q = max(q, int(lines[-1][-1]))
n = len(glob.glob('/tmp/sp9/sgf/*')) and sum(open(f).read().count('(;') for f in glob.glob('/tmp/sp9/sgf/*'))
print("$1 9x9 10bx128 -p 400: %d games in %.1f s -> %.1f games/hour, %.0f NN evals/s (1 GPU, %d host cores, $PG parallel games)" % (n, t, n * 3600 / t, q / t, os.cpu_count()))
PY
}
{
run "sharded-cache" oracle/_ref/sayuri_b200_frontend
run "reference-cache" oracle/_ref/sayuri_b200_frontend_refcache
run "sharded-cache(2)" oracle/_ref/sayuri_b200_frontend
} | tee gpurun_out/selfplay_cache_ab.log
{
echo "| threads | probes per leaf | reference cache, evals/s | sharded cache, evals/s |"
echo "|---|---|---|---|"
for T in 1 16 128 1024; do for P in 1 8; do
  A=$(oracle/_ref/cache_harness_ref bench $T 1.5 400000 2000000 $P | python -c "import json,sys; print('%.0f' % json.loads(sys.stdin.read())['evals_per_s'])")
  B=$(oracle/_ref/cache_harness_b200 bench $T 1.5 400000 2000000 $P | python -c "import json,sys; print('%.0f' % json.loads(sys.stdin.read())['evals_per_s'])")
  echo "| $T | $P | $A | $B |"
done; done
} | tee gpurun_out/cache_bench.md
