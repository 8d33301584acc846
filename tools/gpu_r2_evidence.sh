#!/bin/bash
# Round-2 evidence on one B200: ncu launch lists + one full capture per rung of the current kernels, both bench arms,
# fp16 rung bench lines, the reference GPU comparator built natively and AS SHIPPED (PTX JIT from compute_90).
mkdir -p gpurun_out
for PREC in fp32_split fp16; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_$PREC.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eval-threads 0 --precision $PREC > gpurun_out/r02_ncu_launches_$PREC.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc2 -s 7 -c 1 -f -o gpurun_out/r02_prof_conv_$PREC \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eval-threads 0 --precision $PREC > gpurun_out/r02_ncu_full_$PREC.log 2>&1
done
timeout 300 python bench.py --steps 50 --warmup 5 --precision fp16 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02_bench_fp16.json
timeout 300 python bench.py --steps 30 --warmup 5 --net 20bx256 --no-cpu-baseline --eval-threads 0 2>/dev/null | tail -1 > gpurun_out/r02_bench_split_20bx256.json
timeout 300 python bench.py --steps 30 --warmup 5 --net 20bx256 --precision fp16 --no-cpu-baseline --eval-threads 0 2>/dev/null | tail -1 > gpurun_out/r02_bench_fp16_20bx256.json
for B in sayuri_cudnn_bench sayuri_cudnn_bench_ptx90; do
  SAYURI_CUDNN_BENCH=$B timeout 600 python tools/cudnn_compare.py --net 20bx256 --batches 1,16,256,2048 --seconds 1.0 > gpurun_out/r02_cudnn_compare_20bx256_$B.md 2> gpurun_out/r02_cudnn_compare_20bx256_$B.err
  tail -8 gpurun_out/r02_cudnn_compare_20bx256_$B.md; tail -3 gpurun_out/r02_cudnn_compare_20bx256_$B.err
done
