#!/bin/bash
# Self-play rate of the UNMODIFIED reference loop over our pipe with the host-side replacements linked in
# (sharded NN cache + link-time Board::ComputePassAliveArea): tools/selfplay_host.sh <board> <games> <timeout_s> [frontend]
mkdir -p gpurun_out
[ -f /tmp/fe_10bx128.bin ] || python -c "
import sys; sys.path.insert(0,'.')
from sayuri_b200 import synth
synth.write_synth_net('/tmp/fe_10bx128.bin', '10bx128', seed=20260417)"
BS=${1:-19}; NG=${2:-64}; TO=${3:-200}; FE=${4:-oracle/_ref/sayuri_b200_frontend}
rm -rf /tmp/sph && mkdir -p /tmp/sph
S=$(date +%s.%N)
timeout $TO $FE --mode selfplay -w /tmp/fe_10bx128.bin --no-fp16 -g 0 --parallel-games $NG --num-games $NG -p 400 \
   --selfplay-query bkp:$BS:7:1.0 --target-directory /tmp/sph --cache-memory-mib 2000 2>&1 | tail -1
E=$(date +%s.%N)
python - <<PY
import glob, os
t = $E - $S
q = 0
for f in glob.glob('/tmp/sph/net_queries/*.txt'):
    lines = [x.split() for x in open(f).read().strip().splitlines() if x.strip()]
    if lines: q = max(q, int(lines[-1][-1]))
n = sum(open(f).read().count('(;') for f in glob.glob('/tmp/sph/sgf/*'))
print("$(basename $FE) ${BS}x${BS} 10bx128 -p 400: %d games in %.1f s -> %.1f games/hour, %.0f NN evals/s (1 GPU, %d host cores, $NG parallel games)" % (n, t, n * 3600 / t, q / t, os.cpu_count()))
PY
