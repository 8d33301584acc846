#!/bin/bash
# games/hour of the UNMODIFIED reference self-play loop over our pipe (19x19, 10bx128, 400 visits), plus the
# reference Eigen CPU loop on 9x9 where it finishes in reasonable time.
mkdir -p gpurun_out
python -c "
import sys; sys.path.insert(0,'.')
from sayuri_b200 import synth
synth.write_synth_net('/tmp/fe_10bx128.bin', '10bx128', seed=20260417)
synth.write_synth_net('/tmp/fe_6bx96.bin', '6bx96', seed=20260417)"
NG=${1:-64}
rm -rf /tmp/sp19 && mkdir -p /tmp/sp19
echo "== ours: 19x19 10bx128 -p 400, $NG parallel games, $NG games, fp32-split"
S=$(date +%s.%N)
timeout 1500 oracle/_ref/sayuri_b200_frontend --mode selfplay -w /tmp/fe_10bx128.bin --no-fp16 -g 0 --parallel-games $NG --num-games $NG -p 400 \
   --selfplay-query bkp:19:7:1.0 --target-directory /tmp/sp19 --cache-memory-mib 2000 2>&1 | tail -2
E=$(date +%s.%N)
python - <<PY
import glob
t = $E - $S
q = 0
for f in glob.glob('/tmp/sp19/net_queries/*.txt'):
    q += sum(int(x.split()[-1]) for x in open(f).read().strip().splitlines() if x.strip()) if False else 0
import os
ng = $NG
print("ours_19x19: %d games in %.1f s -> %.1f games/hour (1 GPU, %d host cores)" % (ng, t, ng * 3600 / t, os.cpu_count()))
print(open(glob.glob('/tmp/sp19/net_queries/*.txt')[0]).read()[:400])
PY
rm -rf /tmp/sp9e /tmp/sp9o && mkdir -p /tmp/sp9e /tmp/sp9o
NC=$(nproc)
echo "== reference Eigen: 9x9 6bx96 -p 100, $NC parallel games (config 1 shape)"
S=$(date +%s.%N)
timeout 900 oracle/_ref/sayuri_eigen_v3 --mode selfplay -w /tmp/fe_6bx96.bin --parallel-games $NC --num-games $NC -p 100 -t 1 \
   --selfplay-query bkp:9:7:1.0 --target-directory /tmp/sp9e 2>&1 | tail -1
E=$(date +%s.%N)
python -c "print('eigen_9x9: %d games in %.1f s -> %.1f games/hour' % ($NC, $E-$S, $NC*3600/($E-$S)))"
echo "== ours: 9x9 6bx96 -p 100, 64 parallel games"
S=$(date +%s.%N)
timeout 900 oracle/_ref/sayuri_b200_frontend --mode selfplay -w /tmp/fe_6bx96.bin --no-fp16 -g 0 --parallel-games 64 --num-games 64 -p 100 \
   --selfplay-query bkp:9:7:1.0 --target-directory /tmp/sp9o 2>&1 | tail -1
E=$(date +%s.%N)
python -c "print('ours_9x9: %d games in %.1f s -> %.1f games/hour' % (64, $E-$S, 64*3600/($E-$S)))"
