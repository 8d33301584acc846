#!/bin/bash
# Round-end evidence run: parity tests, both bench arms, launch lists + full ncu captures of the conv kernel on both
# rungs, fixed-seed visit parity through the reference front-end, wide-N timing of the 192-wide net.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader | tee gpurun_out/final_gpu.txt; nproc | tee -a gpurun_out/final_gpu.txt
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/final_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/final_smoke.log
echo "== bench (default line, with cpu baseline)"; timeout 900 python bench.py --steps 50 --warmup 5 2>&1 | tail -1 | tee gpurun_out/final_bench.json | cut -c1-400
echo "== bench --impl reference"; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee gpurun_out/final_bench_reference.json | cut -c1-400
echo "== bench fp16"; timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --precision fp16 2>&1 | tail -1 | tee gpurun_out/final_bench_fp16.json | cut -c1-300
echo "== bench fp16 20bx256 / 15bx192"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision fp16 --net 20bx256 --eval-threads 0 2>&1 | tail -1 | tee gpurun_out/final_bench_fp16_20bx256.json | cut -c1-260
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision fp16 --net 15bx192 --eval-threads 0 2>&1 | tail -1 | tee gpurun_out/final_bench_fp16_15bx192.json | cut -c1-260
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --net 20bx256 --eval-threads 0 2>&1 | tail -1 | tee gpurun_out/final_bench_split_20bx256.json | cut -c1-260
for PREC in fp32_split fp16; do
  echo "== ncu launch list $PREC"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/final_launches_$PREC.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eval-threads 0 --precision $PREC > gpurun_out/ncu_bench.log 2>&1
  python - <<PY
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/final_launches_$PREC.csv')) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows[-31:]:
    k = r[4].split('(')[0][:60]
    agg.setdefault(k, []).append(float(r[14]) / 1e3)
tot = sum(sum(v) for v in agg.values())
for k, v in agg.items():
    print("%-62s n=%2d  mean %.1f us  total %.1f us  share %.1f%%" % (k, len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))
PY
  echo "== ncu full $PREC"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc2 -s 9 -c 1 -o gpurun_out/final_prof_conv_$PREC python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eval-threads 0 --precision $PREC > gpurun_out/ncu_full_$PREC.log 2>&1
  tail -1 gpurun_out/ncu_full_$PREC.log | cut -c1-200
done
echo "== visit parity 19x19 10bx128 (fixed seed, engine batcher in the det front-end)"
timeout 900 python tools/visit_parity.py --net 10bx128 --board 19 --playouts 400 --moves 3 --seeds 1,2 2>&1 | tail -1 | tee gpurun_out/final_visit_parity_19.log
timeout 600 python tools/visit_parity.py --net 6bx96 --board 9 --playouts 400 --moves 6 --seeds 1,2,3 2>&1 | tail -1 | tee gpurun_out/final_visit_parity_9.log
