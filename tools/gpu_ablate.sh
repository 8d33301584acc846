#!/bin/bash
# Ablations of conv3x3_tc2 (timing only; results are wrong by construction when a bit is set):
#  8 = zero tap shifts (8-row aligned A operand), 16 = no weight TMA, 32 = half N per MMA, 64 = no slab TMA,
#  1 = no epilogue stores / residual loads, 2 = no activation math.
mkdir -p gpurun_out
for prec in 0 1; do
  for dbg in 0 8 16 24 32 64 88 1 3 91; do
    timeout 120 python tools/conv_stats.py --precision $prec --dbg $dbg --launch 2 2>&1 | grep -E "^net|mma_total|wait_|epi_total|per item"
  done
done | tee gpurun_out/ablate.log
