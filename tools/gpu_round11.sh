#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --eval-threads 0 2>&1 | tail -1 | tee gpurun_out/bench_r11.json | cut -c1-300
echo "== bench fp16"; timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --precision fp16 --eval-threads 0 2>&1 | tail -1 | tee gpurun_out/bench_fp16_r11.json | cut -c1-300
python - <<'PY'
import json
for f in ("bench_r11", "bench_fp16_r11"):
    d = json.load(open("gpurun_out/%s.json" % f)); r = d["roofline"]
    print(f, "value %.0f ms/step %.4f conv_ms %.4f share %.3f frac %.4f launches %d" % (d["value"], d["ms_per_step"], r["kernel_ms_per_step"], r["kernel_share_of_step"], r["frac"], r["launches_per_step"]))
PY
for PREC in fp32_split fp16; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r11_$PREC.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eval-threads 0 --precision $PREC > gpurun_out/ncu_bench.log 2>&1
  python - <<PY
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_r11_$PREC.csv')) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows[-31:]:
    k = r[4].split('(')[0][:60]
    agg.setdefault(k, []).append(float(r[14]) / 1e3)
for k, v in agg.items():
    print("$PREC %-50s n=%2d  mean %.1f us" % (k, len(v), sum(v) / len(v)))
PY
done
