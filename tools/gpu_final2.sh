#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/final_pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/final_smoke.log
echo "== bench"; timeout 900 python bench.py --steps 50 --warmup 5 2>&1 | tail -1 | tee gpurun_out/final_bench.json | cut -c1-200
echo "== bench 20bx256"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --net 20bx256 --eval-threads 0 2>&1 | tail -1 | tee gpurun_out/final_bench_split_20bx256.json | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/final_launches_fp32_split.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eval-threads 0 > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc2 -s 9 -c 1 -o gpurun_out/final_prof_conv_fp32_split python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eval-threads 0 > gpurun_out/ncu_full_fp32_split.log 2>&1
tail -1 gpurun_out/ncu_full_fp32_split.log | cut -c1-160
python - <<'PY'
import json
for f in ("final_bench", "final_bench_split_20bx256"):
    d = json.load(open("gpurun_out/%s.json" % f)); r = d["roofline"]
    print(f, "value %.0f e2e %.0f frac %.4f share %.3f clocks %s" % (d["value"], d["e2e"]["value"], r["frac"], r["kernel_share_of_step"], d["clocks"]["sm_mhz"]))
PY
