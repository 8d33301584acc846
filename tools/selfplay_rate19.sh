#!/bin/bash
# games/hour of the UNMODIFIED reference self-play loop over our pipe (19x19, 10bx128, 400 visits): engine batcher vs
# the reference batcher (SAYURI_B200_REF_BATCHER=1), same weights, same options.
mkdir -p gpurun_out
python -c "
import sys; sys.path.insert(0,'.')
from sayuri_b200 import synth
synth.write_synth_net('/tmp/fe_10bx128.bin', '10bx128', seed=20260417)"
NG=${1:-64}
run() {
  rm -rf /tmp/sp19 && mkdir -p /tmp/sp19
  S=$(date +%s.%N)
  env $2 timeout 1200 oracle/_ref/sayuri_b200_frontend --mode selfplay -w /tmp/fe_10bx128.bin --no-fp16 -g 0 --parallel-games $NG --num-games $NG -p 400 \
     --selfplay-query bkp:19:7:1.0 --target-directory /tmp/sp19 --cache-memory-mib 2000 2>&1 | tail -1
  E=$(date +%s.%N)
  python - <<PY
import glob, os
t = $E - $S
ng = $NG
q = 0
for f in glob.glob('/tmp/sp19/net_queries/*.txt'):
    lines = [x.split() for x in open(f).read().strip().splitlines() if x.strip()]
    if lines: q = max(q, int(lines[-1][-1]))
print("$1 19x19 10bx128 -p 400: %d games in %.1f s -> %.1f games/hour, %.0f NN evals/s (1 GPU, %d host cores, %d parallel games)" % (ng, t, ng * 3600 / t, q / t, os.cpu_count(), ng))
PY
}
run "engine-batcher" "SAYURI_B200_REF_BATCHER=0" | tee gpurun_out/selfplay19.log
if [ "${2:-both}" = "both" ]; then run "reference-batcher" "SAYURI_B200_REF_BATCHER=1" | tee -a gpurun_out/selfplay19.log; fi
