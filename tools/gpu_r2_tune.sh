#!/bin/bash
# Round 2: small-batch latency with layer_overlap on/off, weight-ring / slab-buffer variants of the split rung, one full bench line.
mkdir -p gpurun_out
timeout 300 python tools/latency_sweep.py --option layer_overlap 2>&1 | tee gpurun_out/r2_latency_overlap.md
one() {
  env $3 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --eval-threads 0 --option layer_overlap=0 $2 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 value %.0f evals/s ms/step %.4f conv_ms %.4f conv_share %.3f frac %.4f clocks %s %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['kernel_share_of_step'], d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons']))" | tee -a gpurun_out/r2_tune_ab.log
}
: > gpurun_out/r2_tune_ab.log
for i in 1 2; do
  one "split 2 slabs 12 stages" "" X=1
  one "split 2 slabs 16 stages" "" SAYURI_B200_LIB=build/libsb_nb16.so
  one "split 3 slabs  6 stages" "" SAYURI_B200_LIB=build/libsb_na3nb6.so
  one "fp16                   " "--precision fp16" X=1
done
one "20bx256 split 12 stages" "--net 20bx256 --steps 20" X=1
one "20bx256 split 16 stages" "--net 20bx256 --steps 20" SAYURI_B200_LIB=build/libsb_nb16.so
one "20bx256 fp16           " "--net 20bx256 --steps 20 --precision fp16" X=1
one "15bx192 split          " "--net 15bx192 --steps 20" X=1
one "15bx192 fp16           " "--net 15bx192 --steps 20 --precision fp16" X=1
timeout 300 python bench.py --steps 50 --warmup 5 --option layer_overlap=0 2>gpurun_out/r2_tune_bench.err | tail -1 > gpurun_out/r2_tune_bench.json; cut -c1-400 gpurun_out/r2_tune_bench.json
