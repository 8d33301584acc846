#!/bin/bash
# one tower conv launch (mish, split rung, batch 256) of the round-1 library and of the current one under ncu --set full
mkdir -p gpurun_out
for which in old new; do
  if [ $which = old ]; then export SAYURI_B200_LIB=build/libsb_old.so; else unset SAYURI_B200_LIB; fi
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc2 -s 7 -c 1 -f -o gpurun_out/r2_prof_conv_$which \
     python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eval-threads 0 > gpurun_out/r2_ncu_$which.log 2>&1
  tail -1 gpurun_out/r2_ncu_$which.log | cut -c1-150
done
