#!/bin/bash
# A/B of the layer3 conv3 + next-conv1 seam kernel (HMV_FUSE_NEXT=1): bisect vs oracle, then bench with / without
mkdir -p gpurun_out
HMV_FUSE_NEXT=1 timeout 200 python tools/diag.py model_bf16 > gpurun_out/diag_seam.log 2>&1; echo "diag rc $?"
grep -E "layer3|e2e|EXCEPTION|rror" gpurun_out/diag_model_bf16.txt | head -40
for v in 1 0; do
  HMV_FUSE_NEXT=$v timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-clocks > gpurun_out/bench_seam$v.json 2>gpurun_out/bench_seam$v.err; echo "bench $v rc $?"
  cp gpurun_out/tc_launches.csv gpurun_out/tc_launches_seam$v.csv
  python -c "
import json
d=json.load(open('gpurun_out/bench_seam$v.json'))
print('FUSE_NEXT=$v value %.0f ms %.2f median %.2f'%(d['value'], d['ms_per_step'], d['step_ms']['median']), {k: round(x,3) for k,x in d['roofline']['phase_ms_per_step'].items()})"
done
grep -E "layer3.1" gpurun_out/tc_launches_seam1.csv | head -4; grep -E "layer3.1" gpurun_out/tc_launches_seam0.csv | head -4
tail -3 gpurun_out/bench_seam1.err
