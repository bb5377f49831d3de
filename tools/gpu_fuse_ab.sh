#!/bin/bash
# A/B of the fused bottleneck tail: bisect vs oracle, GPU tests, bench with / without the fusion
mkdir -p gpurun_out
HMV_FUSE_TAIL=7 timeout 300 python tools/diag.py model_bf16 > gpurun_out/diag_model_bf16.log 2>&1
echo "diag exit $?" > gpurun_out/phases.txt
grep -E "rel-L2|EXCEPTION|Error|error" gpurun_out/diag_model_bf16.txt | awk '{ if ($0 ~ /rel-L2/) print }' | head -70
tail -5 gpurun_out/diag_model_bf16.log
timeout 1500 python -m pytest tests -q -m gpu --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/phases.txt
tail -3 gpurun_out/pytest_gpu.log
grep -E "^(FAILED|E  )|teacher-forced|flip rate" gpurun_out/pytest_gpu.log | cut -c1-250
for mode in fused unfused; do
  if [ $mode == unfused ]; then export HMV_FUSE_TAIL=0; else export HMV_FUSE_TAIL=${FUSE_MASK:-7}; fi
  timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$mode.json 2>> gpurun_out/bench.err
  echo "bench $mode exit $?" >> gpurun_out/phases.txt
  cp gpurun_out/tc_launches.csv gpurun_out/tc_launches_$mode.csv
done
unset HMV_FUSE_TAIL
cat gpurun_out/phases.txt; for f in gpurun_out/bench_fused.json gpurun_out/bench_unfused.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print('value %.0f e2e %.0f ms %.2f tc_frac %.3f share %.2f tc_ms %.2f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['roofline']['kernel_ms_per_step']), {k: round(v,2) for k,v in d['step_ms'].items()}, d['roofline']['phase_ms_per_step'])
"; done
grep -E "conv2|conv3" gpurun_out/tc_launches_fused.csv | head -13
HMV_FUSE_TAIL=7 HMV_BT_PROF=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-clocks --ramp-seconds 0 2>&1 | grep bt_prof | head -13
