#!/bin/bash
# default seam kernel still correct after the template change? opt-in cluster variant correct? A/B timing.
mkdir -p gpurun_out
timeout 60 python tools/diag.py model_bf16 > gpurun_out/diag_a.log 2>&1; echo "diag default rc $?"; grep -E "layer3.5.conv3|e2e feat" gpurun_out/diag_model_bf16.txt
HMV_SEAM_CLUSTER=1 timeout 60 python tools/diag.py model_bf16 > gpurun_out/diag_b.log 2>&1; echo "diag cluster rc $?"; grep -E "layer3.[0-5].conv3|e2e feat|EXCEPTION" gpurun_out/diag_model_bf16.txt
for v in 1 0; do
HMV_SEAM_CLUSTER=$v timeout 60 python bench.py --steps 12 --warmup 3 --ramp-seconds 0.5 --no-cpu-baseline --no-e2e --no-clocks 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('SEAM_CLUSTER=$v value %.0f ms %.2f median %.2f'%(d['value'], d['ms_per_step'], d['step_ms']['median']))"
grep -E "layer3.2.conv3" gpurun_out/tc_launches.csv | head -1
done
