#!/bin/bash
# diag + tests + bench on the GPU box
mkdir -p gpurun_out
timeout 300 python tools/diag.py model_bf16 > gpurun_out/diag_model_bf16.log 2>&1
echo "diag exit $?" > gpurun_out/phases.txt
head -12 gpurun_out/diag_model_bf16.txt; tail -7 gpurun_out/diag_model_bf16.txt
timeout 1500 python -m pytest tests -q -m gpu --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/phases.txt
tail -3 gpurun_out/pytest_gpu.log
grep -E "^(FAILED|E  )|teacher-forced|flip rate" gpurun_out/pytest_gpu.log | cut -c1-250
for mb in 32 64; do
  for rep in a b; do
  timeout 600 python bench.py --steps 40 --warmup 5 --micro-batch $mb --no-cpu-baseline > gpurun_out/bench_mb${mb}_$rep.json 2>> gpurun_out/bench.err
  echo "bench mb$mb $rep exit $?" >> gpurun_out/phases.txt
  done
done
if [ "$1" == "ncu" ]; then
timeout 600 python bench.py --steps 2 --warmup 3 --ramp-seconds 0 --micro-batch 64 --no-cpu-baseline --no-clocks > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 100 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --ramp-seconds 0 --micro-batch 64 --no-cpu-baseline --no-clocks > gpurun_out/ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/phases.txt
fi
cat gpurun_out/phases.txt; for f in gpurun_out/bench_mb*_?.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print('value %.0f e2e %.0f ms %.2f tc_frac %.3f share %.2f tc_ms %.2f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['roofline']['kernel_ms_per_step']), {k: round(v,2) for k,v in d['step_ms'].items()}, d['roofline']['phase_ms_per_step'])
"; done
