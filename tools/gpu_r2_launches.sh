#!/bin/bash
# launch lists (every kernel with its device time, serialised) of one forward: round-1 library vs current
mkdir -p gpurun_out
for which in "$@"; do
  label=${which%%=*}; lib=${which#*=}
  if [ "$lib" = "-" ]; then unset SAYURI_B200_LIB; else export SAYURI_B200_LIB=$lib; fi
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_$label.csv \
     python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eval-threads 0 > gpurun_out/r2_launches_$label.log 2>&1
done
