#!/bin/bash
mkdir -p gpurun_out
{
echo "== old"; SAYURI_B200_LIB=build/libsb_old.so python tools/conv_stats.py
echo "== new chunk 9"; python tools/conv_stats.py
echo "== new chunk 0"; python tools/conv_stats.py --option chunk_taps=0
echo "== new chunk 3"; python tools/conv_stats.py --option chunk_taps=3
} 2>&1 | tee gpurun_out/r2_conv_stats.log
