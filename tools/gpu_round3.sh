#!/bin/bash
mkdir -p gpurun_out
echo "== cluster occupancy"; ./tools/probes/cluster_occupancy | tee gpurun_out/cluster_occupancy.log
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
echo "== pdl on/off"
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/pdl.log
import os, sys, tempfile
sys.path.insert(0, '.')
import numpy as np
from sayuri_b200 import engine, synth
for net, b in (("10bx128", 256), ("10bx128", 32), ("20bx256", 256), ("20bx256", 8)):
    path = os.path.join(tempfile.gettempdir(), "pdl_%s.bin" % net); synth.write_synth_net(path, net, seed=20260417)
    pos = synth.synth_positions(64, 19, seed=5).reshape(64, -1)
    for prec in (0, 1):
        pipe = engine.B200ForwardPipe().initialize(path, 19, b, gpus=[0], precision=prec)
        planes = [pos[i % 64] for i in range(b)]
        pipe.batch_forward(0, planes, [19]*b, [0]*b)
        for pdl in (0, 1, 0, 1):
            pipe.set_option("pdl", pdl)
            pipe.time_forward(0, 0, 5, flush_l2=True)
            ms, _, _ = pipe.time_forward(0, 0, 30, flush_l2=True)
            print("%s batch %d precision %d pdl %d: %.4f ms median, %.0f evals/s" % (net, b, prec, pdl, float(np.median(ms)), b / float(np.median(ms)) * 1e3), flush=True)
        pipe.destroy()
PY
echo "== bench"; timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1200 | tee gpurun_out/bench.log
echo "== ncu launch list (split)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_split.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eval-threads 0 > gpurun_out/ncu_bench.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_split.csv')) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows[-31:]:
    k = r[4].split('(')[0][:60]
    agg.setdefault(k, []).append(float(r[14]) / 1e3)
for k, v in agg.items():
    print("%-62s n=%2d  mean %.1f us  total %.1f us" % (k, len(v), sum(v) / len(v), sum(v)))
PY
