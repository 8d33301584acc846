#!/bin/bash
# Round 2: the whole GPU test suite on the chained build, then bench against the library before the chain work.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --durations=5 2>&1 | tail -12 | tee gpurun_out/r2_prefinal_pytest.log
one() {
  env $3 timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --eval-threads 0 $2 2>gpurun_out/r2_prefinal_last.err | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1 value %.0f evals/s ms/step %.4f conv_ms %.4f conv_share %.3f frac %.4f launches %d clocks %s %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['kernel_share_of_step'], d['roofline']['frac'], d['gpu_launches'], d['clocks']['sm_mhz'], d['clocks']['reasons']))" | tee -a gpurun_out/r2_prefinal_ab.log
}
: > gpurun_out/r2_prefinal_ab.log
for i in 1 2; do
  one "split prev lib " "" SAYURI_B200_LIB=build/libsb_prev.so
  one "split chain    " "" X=1
  one "split chain off" "--option conv_chain=0" X=1
  one "fp16  prev lib " "--precision fp16" SAYURI_B200_LIB=build/libsb_prev.so
  one "fp16  chain    " "--precision fp16" X=1
  one "fp16  chain off" "--precision fp16 --option conv_chain=0" X=1
done
