#!/bin/bash
# Round 2: the unrolled single-thread MMA issuer against the general issue loop (conv_dbg 128), same library, same box.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "scheduling_knobs or golden_vectors or every_activation or channel_widths" 2>&1 | tail -5 | tee gpurun_out/r2_issuer_pytest.log
{
for PREC in 0 1; do
  for DBG in 0 128; do
    echo "== precision $PREC conv_dbg $DBG"; timeout 200 python tools/conv_stats.py --precision $PREC --dbg $DBG
  done
done
echo "== 20bx256 split"; timeout 200 python tools/conv_stats.py --net 20bx256 --dbg 0
echo "== 20bx256 split general"; timeout 200 python tools/conv_stats.py --net 20bx256 --dbg 128
} 2>&1 | tee gpurun_out/r2_issuer_stats.log
one() {
  timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --eval-threads 0 $2 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 value %.0f evals/s ms/step %.4f conv_ms %.4f conv_share %.3f frac %.4f clocks %s %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['kernel_share_of_step'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons']))" | tee -a gpurun_out/r2_issuer_ab.log
}
: > gpurun_out/r2_issuer_ab.log
for i in 1 2; do
  one "split unrolled" ""
  one "split general " "--option conv_dbg=128"
  one "fp16  unrolled" "--precision fp16"
  one "fp16  general " "--precision fp16 --option conv_dbg=128"
done
