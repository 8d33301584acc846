#!/bin/bash
mkdir -p gpurun_out
PREC=${1:-fp32_split}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$PREC.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --precision $PREC > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc2 -s 7 -c 1 -o gpurun_out/prof_conv_$PREC python bench.py --steps 1 --warmup 3 --no-cpu-baseline --precision $PREC > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
