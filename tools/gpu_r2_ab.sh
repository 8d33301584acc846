#!/bin/bash
# same-box A/B of library builds, alternating, config 2 (split rung): tools/gpu_r2_ab.sh "<label>=<lib or ->[:option]" ...
mkdir -p gpurun_out
: > gpurun_out/r2_ab.log
one() {
  env $1 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --eval-threads 0 $3 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$2 value %.0f evals/s ms/step %.4f conv_ms %.4f conv_share %.3f clocks %s %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['kernel_share_of_step'], d['clocks']['sm_mhz'], d['clocks']['reasons']))" | tee -a gpurun_out/r2_ab.log
}
for i in 1 2 3; do
  for spec in "$@"; do
    label=${spec%%=*}; rest=${spec#*=}; lib=${rest%%:*}; opt=""
    if [[ "$rest" == *:* ]]; then opt="--option ${rest#*:}"; fi
    if [ "$lib" = "-" ]; then one "X=1" "$label" "$opt"; else one "SAYURI_B200_LIB=$lib" "$label" "$opt"; fi
  done
done
