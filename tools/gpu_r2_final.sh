#!/bin/bash
# Round 2, final evidence on one B200: whole GPU test suite, smoke, both bench arms, other nets / rungs, ncu launch lists and
# full captures of the conv kernel (one layer per launch and a chained launch), in-kernel counters.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --durations=5 2>&1 | tail -12 | tee gpurun_out/r02f_pytest_gpu.log
cp /tmp/sb_visit_parity.log gpurun_out/r02f_visit_parity.log 2>/dev/null
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/r02f_smoke.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/r02f_bench_reference.json
timeout 300 python bench.py --steps 50 --warmup 5 2>gpurun_out/r02f_bench.err | tail -1 > gpurun_out/r02f_bench.json; cut -c1-300 gpurun_out/r02f_bench.json
timeout 300 python bench.py --steps 50 --warmup 5 --precision fp16 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02f_bench_fp16.json
timeout 300 python bench.py --steps 30 --warmup 5 --net 20bx256 --no-cpu-baseline --eval-threads 0 2>/dev/null | tail -1 > gpurun_out/r02f_bench_split_20bx256.json
timeout 300 python bench.py --steps 30 --warmup 5 --net 20bx256 --precision fp16 --no-cpu-baseline --eval-threads 0 2>/dev/null | tail -1 > gpurun_out/r02f_bench_fp16_20bx256.json
timeout 300 python bench.py --steps 30 --warmup 5 --net 15bx192 --precision fp16 --no-cpu-baseline --eval-threads 0 2>/dev/null | tail -1 > gpurun_out/r02f_bench_fp16_15bx192.json
for PREC in fp32_split fp16; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02f_launches_$PREC.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --eval-threads 0 --precision $PREC > gpurun_out/r02f_ncu_launches_$PREC.log 2>&1
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc2 -s 7 -c 1 -f -o gpurun_out/r02f_prof_conv_$PREC \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eval-threads 0 --precision $PREC --option conv_chain=0 > gpurun_out/r02f_ncu_full_$PREC.log 2>&1
  timeout 300 ncu --set full --clock-control none -k regex:conv3x3_tc2 -s 4 -c 1 -f -o gpurun_out/r02f_prof_chain_$PREC \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --eval-threads 0 --precision $PREC > gpurun_out/r02f_ncu_chain_$PREC.log 2>&1
done
{
for PREC in 0 1; do echo "== precision $PREC"; timeout 200 python tools/conv_stats.py --precision $PREC; done
} 2>&1 | tee gpurun_out/r02f_conv_stats.log
