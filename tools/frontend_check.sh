#!/bin/bash
# The UNMODIFIED reference front-end (GTP / netbench / self-play) over our pipe, on the GPU box.
mkdir -p gpurun_out
W=/tmp/fe_10bx128.bin
python -c "
import sys; sys.path.insert(0,'.')
from sayuri_b200 import synth
synth.write_synth_net('$W', '10bx128', seed=20260417)
synth.write_synth_net('/tmp/fe_6bx96.bin', '6bx96', seed=20260417)"
FE=oracle/_ref/sayuri_b200_frontend
echo "== netbench (reference's own GTP tool) over sayuri_b200, fp32-split"
printf 'netbench timelimit 5 batchsize 32 256\nquit\n' | timeout 200 $FE -w $W --no-fp16 -g 0 -b 256 2>&1 | grep -E "batch size=|Backend|sayuri_b200|rror" | tee gpurun_out/netbench_b200.log
echo "== netbench over sayuri_b200, fp16"
printf 'netbench timelimit 5 batchsize 256\nquit\n' | timeout 200 $FE -w $W -g 0 -b 256 2>&1 | grep -E "batch size=|rror" | tee -a gpurun_out/netbench_b200.log
echo "== netbench reference Eigen (single thread by construction)"
printf 'netbench timelimit 5\nquit\n' | timeout 200 oracle/_ref/sayuri_eigen_v3 -w $W 2>&1 | grep -E "batch size=" | tee gpurun_out/netbench_eigen.log
echo "== self-play plumbing 9x9 (config 1 shape: 6bx96, 100 visits) over sayuri_b200"
rm -rf /tmp/sp9 && mkdir -p /tmp/sp9
( time timeout 600 $FE --mode selfplay -w /tmp/fe_6bx96.bin --no-fp16 -g 0 --parallel-games 16 --num-games 16 -p 100 \
   --selfplay-query bkp:9:7:1.0 --target-directory /tmp/sp9 --cache-memory-mib 400 ) 2>&1 | tail -6 | tee gpurun_out/selfplay9.log
ls /tmp/sp9 /tmp/sp9/* 2>/dev/null | head -12 | tee -a gpurun_out/selfplay9.log
echo "== visit parity 9x9"
timeout 900 python tools/visit_parity.py --net 6bx96 --board 9 --playouts 400 --moves 6 --seeds 1,2,3 2>&1 | tail -2 | tee gpurun_out/visit_parity.log
