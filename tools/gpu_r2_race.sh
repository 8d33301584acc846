#!/bin/bash
# Round 2: race hunt for the chained / overlapped convolution launches (bit-exact against the plain launches), two ways of
# releasing the per-tile counters; then a bench of each.
mkdir -p gpurun_out
{
for LIB in sayuri_b200/libsayuri_b200.so build/libsb_perwarp.so; do
  for PREC in 0 1; do
    echo "== $LIB precision $PREC conv_chain=1"; SAYURI_B200_LIB=$LIB timeout 120 python tools/chain_race.py --precision $PREC --option conv_chain=1 2>&1 | tail -6
    echo "== $LIB precision $PREC layer_overlap=2"; SAYURI_B200_LIB=$LIB timeout 120 python tools/chain_race.py --precision $PREC --option layer_overlap=2 2>&1 | tail -6
  done
done
echo "== build/libsb_prev.so precision 0 layer_overlap=2"; SAYURI_B200_LIB=build/libsb_prev.so timeout 120 python tools/chain_race.py --precision 0 --option layer_overlap=2 2>&1 | tail -6
echo "== 20bx256 default lib conv_chain=1"; timeout 200 python tools/chain_race.py --net 20bx256 --repeats 3 --option conv_chain=1 2>&1 | tail -4
} | tee gpurun_out/r2_race.log
one() {
  env $3 timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --eval-threads 0 $2 2>gpurun_out/r2_race_last.err | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 value %.0f evals/s ms/step %.4f conv_ms %.4f conv_share %.3f frac %.4f launches %d clocks %s %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['kernel_share_of_step'], d['roofline']['frac'], d['gpu_launches'], d['clocks']['sm_mhz'], d['clocks']['reasons']))" | tee -a gpurun_out/r2_race_ab.log
}
: > gpurun_out/r2_race_ab.log
for i in 1 2; do
  one "split prev lib       " "" SAYURI_B200_LIB=build/libsb_prev.so
  one "split chain publisher" "" X=1
  one "split chain per-warp " "" SAYURI_B200_LIB=build/libsb_perwarp.so
  one "fp16  prev lib       " "--precision fp16" SAYURI_B200_LIB=build/libsb_prev.so
  one "fp16  chain publisher" "--precision fp16" X=1
  one "fp16  chain per-warp " "--precision fp16" SAYURI_B200_LIB=build/libsb_perwarp.so
done
