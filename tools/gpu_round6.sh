#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-1500 | tee gpurun_out/bench.log
echo "== bench fp16"; timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --precision fp16 2>&1 | tail -1 | cut -c1-600 | tee gpurun_out/bench_fp16.log
for PREC in fp32_split fp16; do
  echo "== ncu launch list $PREC"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$PREC.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eval-threads 0 --precision $PREC > gpurun_out/ncu_bench.log 2>&1
  python - <<PY
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_$PREC.csv')) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows[-31:]:
    k = r[4].split('(')[0][:60]
    agg.setdefault(k, []).append(float(r[14]) / 1e3)
for k, v in agg.items():
    print("%-62s n=%2d  mean %.1f us  total %.1f us" % (k, len(v), sum(v) / len(v), sum(v)))
PY
  echo "== ncu full $PREC"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc2 -s 9 -c 1 -o gpurun_out/prof_conv_$PREC python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eval-threads 0 --precision $PREC > gpurun_out/ncu_full_$PREC.log 2>&1
  tail -2 gpurun_out/ncu_full_$PREC.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep
