#!/bin/bash
# the whole GPU suite on the final tree, one clean log
mkdir -p gpurun_out
rm -f /tmp/sb_visit_parity.log
timeout 600 python -m pytest tests -q -m gpu --durations=5 2>&1 | tail -10 | tee gpurun_out/r02f_pytest_gpu_final.log
cp /tmp/sb_visit_parity.log gpurun_out/r02f_visit_parity_final.log 2>/dev/null
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r02f_smoke_final.log
