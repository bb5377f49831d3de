#!/bin/bash
mkdir -p gpurun_out
HMV_FUSE_NEXT=1 timeout 200 python tools/diag.py model_bf16 > gpurun_out/diag_seam.log 2>&1; echo "diag rc $?"
grep -E "layer3.[45]|e2e feat|EXCEPTION|rror" gpurun_out/diag_model_bf16.txt | head -8
HMV_FUSE_NEXT=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-clocks > gpurun_out/bench_seam1.json 2>gpurun_out/bench_seam1.err; echo "bench rc $?"
python -c "
import json
d=json.load(open('gpurun_out/bench_seam1.json'))
print('FUSE_NEXT=1 value %.0f ms %.2f median %.2f'%(d['value'], d['ms_per_step'], d['step_ms']['median']), {k: round(x,3) for k,x in d['roofline']['phase_ms_per_step'].items()})"
grep -E "layer3.[12]" gpurun_out/tc_launches.csv | head -5
