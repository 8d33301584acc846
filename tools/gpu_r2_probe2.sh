#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/precision_probe.py 2>&1 | tee gpurun_out/r2_precision_probe3.log
