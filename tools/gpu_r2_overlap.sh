#!/bin/bash
# Round 2: tile-by-tile dependencies between consecutive convolution launches (option layer_overlap) on / off, tail split on / off.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_gpu_fullnets.py -q -m gpu -x 2>&1 | tail -5 | tee gpurun_out/r2_overlap_pytest.log
{
for PREC in 0 1; do
  for OV in 1 0; do
    echo "== precision $PREC layer_overlap $OV tail_split 1"; timeout 200 python tools/conv_stats.py --precision $PREC --option layer_overlap=$OV
  done
  echo "== precision $PREC layer_overlap 1 tail_split 0"; timeout 200 python tools/conv_stats.py --precision $PREC --tail-split 0
done
} 2>&1 | tee gpurun_out/r2_overlap_stats.log
one() {
  timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --eval-threads 0 $2 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 value %.0f evals/s ms/step %.4f conv_ms %.4f conv_share %.3f frac %.4f clocks %s %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['kernel_share_of_step'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons']))" | tee -a gpurun_out/r2_overlap_ab.log
}
: > gpurun_out/r2_overlap_ab.log
for i in 1 2; do
  one "split overlap       " ""
  one "split overlap tail0 " "--option tail_split=0"
  one "split no overlap    " "--option layer_overlap=0"
  one "fp16  overlap       " "--precision fp16"
  one "fp16  overlap tail0 " "--precision fp16 --option tail_split=0"
  one "fp16  no overlap    " "--precision fp16 --option layer_overlap=0"
done
one "20bx256 split overlap   " "--net 20bx256 --steps 20"
one "20bx256 split no overlap" "--net 20bx256 --steps 20 --option layer_overlap=0"
one "20bx256 fp16  overlap   " "--net 20bx256 --steps 20 --precision fp16"
one "20bx256 fp16  no overlap" "--net 20bx256 --steps 20 --precision fp16 --option layer_overlap=0"
one "batch 16 split overlap   " "--batch 16"
one "batch 16 split no overlap" "--batch 16 --option layer_overlap=0"
