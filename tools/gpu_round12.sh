#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench"; SB_PROFILE_GROUPS=1 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --eval-threads 0 2> gpurun_out/groups_split.err | tail -1 | tee gpurun_out/bench_r12.json | cut -c1-200; grep sb_time_forward gpurun_out/groups_split.err
echo "== bench fp16"; SB_PROFILE_GROUPS=1 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --precision fp16 --eval-threads 0 2> gpurun_out/groups_fp16.err | tail -1 | tee gpurun_out/bench_fp16_r12.json | cut -c1-200; grep sb_time_forward gpurun_out/groups_fp16.err
python - <<'PY'
import json
for f in ("bench_r12", "bench_fp16_r12"):
    d = json.load(open("gpurun_out/%s.json" % f)); r = d["roofline"]
    print(f, "value %.0f ms/step %.4f conv_ms %.4f share %.3f frac %.4f launches %d gpu_launches %d" % (d["value"], d["ms_per_step"], r["kernel_ms_per_step"], r["kernel_share_of_step"], r["frac"], r["launches_per_step"], d["gpu_launches"]))
PY
