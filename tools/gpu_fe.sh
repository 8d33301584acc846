#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
W=/tmp/fe_10bx128.bin
python -c "
import sys; sys.path.insert(0,'.')
from sayuri_b200 import synth
synth.write_synth_net('$W', '10bx128', seed=20260417)"
FE=oracle/_ref/sayuri_b200_frontend
echo "== netbench, engine batcher (512 / 64 search threads)"
printf 'netbench timelimit 5 batchsize 32 256\nquit\n' | timeout 200 $FE -w $W --no-fp16 -g 0 -b 256 2>&1 | grep -E "batch size=|rror" | tee gpurun_out/netbench2.log
echo "== bench"; timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_fe.json | cut -c1-200
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_fe.json")); r = d["roofline"]
print("value %.0f e2e %.0f e2e_eval %.0f frac %.4f" % (d["value"], d["e2e"]["value"], d["e2e_eval"]["value"], r["frac"]))
PY
