#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_gpu_fullnets.py tests/test_gpu_engine.py -q 2>&1 | tail -5 | tee gpurun_out/r2_aux_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_aux_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eval-threads 0 > gpurun_out/r2_aux_launches.log 2>&1
for i in 1 2; do timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --eval-threads 0 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.0f ms %.4f share %.3f frac %.3f clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_share_of_step'], d['roofline']['frac'], d['clocks']['sm_mhz']))"; done | tee gpurun_out/r2_aux_bench.log
