#!/bin/bash
# Round 2: 2 vs 4 epilogue warps per TMEM lane quadrant (build/libsb_p4.so) and tail-wave split on/off, now that the MMA
# issuer is no longer the limiter.  Same box, alternating.
mkdir -p gpurun_out
P4=build/libsb_p4.so
SAYURI_B200_LIB=$P4 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/r2_epi_pytest_p4.log
{
for PREC in 0 1; do
  for TS in 1 0; do
    echo "== p2 precision $PREC tail_split $TS"; timeout 200 python tools/conv_stats.py --precision $PREC --tail-split $TS
    echo "== p4 precision $PREC tail_split $TS"; SAYURI_B200_LIB=$P4 timeout 200 python tools/conv_stats.py --precision $PREC --tail-split $TS
  done
done
} 2>&1 | tee gpurun_out/r2_epi_stats.log
one() {
  env $3 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --eval-threads 0 $2 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 value %.0f evals/s ms/step %.4f conv_ms %.4f conv_share %.3f frac %.4f clocks %s %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['kernel_share_of_step'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons']))" | tee -a gpurun_out/r2_epi_ab.log
}
: > gpurun_out/r2_epi_ab.log
for i in 1 2; do
  one "split p2       " "" X=1
  one "split p4       " "" SAYURI_B200_LIB=$P4
  one "split p2 tail0 " "--option tail_split=0" X=1
  one "split p4 tail0 " "--option tail_split=0" SAYURI_B200_LIB=$P4
  one "fp16  p2       " "--precision fp16" X=1
  one "fp16  p4       " "--precision fp16" SAYURI_B200_LIB=$P4
  one "fp16  p4 tail0 " "--precision fp16 --option tail_split=0" SAYURI_B200_LIB=$P4
done
one "20bx256 split p2" "--net 20bx256 --steps 20" X=1
one "20bx256 split p4" "--net 20bx256 --steps 20" SAYURI_B200_LIB=$P4
one "20bx256 fp16  p2" "--net 20bx256 --steps 20 --precision fp16" X=1
one "20bx256 fp16  p4" "--net 20bx256 --steps 20 --precision fp16" SAYURI_B200_LIB=$P4
