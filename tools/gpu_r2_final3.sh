#!/bin/bash
mkdir -p gpurun_out
rm -f /tmp/sb_visit_parity.log
timeout 900 python -m pytest tests/test_gpu_frontend.py -q -m gpu --durations=3 2>&1 | tail -8 | tee gpurun_out/r02f_pytest_frontend.log
cp /tmp/sb_visit_parity.log gpurun_out/r02f_visit_parity.log 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:conv3x3_tc2 -s 3 -c 1 -f -o gpurun_out/r02f_prof_chain_fp16 \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eval-threads 0 --precision fp16 > gpurun_out/r02f_ncu_chain_fp16.log 2>&1
