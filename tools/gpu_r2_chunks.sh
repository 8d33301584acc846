#!/bin/bash
# perf of the chunked-accumulation settings (split rung), config 2 and 20bx256
mkdir -p gpurun_out
: > gpurun_out/r2_chunk_perf.log
for net in 10bx128 20bx256; do
for ct in 0 9 3 1; do
  timeout 300 python bench.py --net $net --steps 40 --warmup 5 --no-cpu-baseline --eval-threads 0 --option chunk_taps=$ct 2>/dev/null | tail -1 | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$net chunk_taps=$ct value %.0f evals/s ms/step %.4f frac %.4f conv_share %.3f clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['clocks']['sm_mhz']))" | tee -a gpurun_out/r2_chunk_perf.log
done; done
