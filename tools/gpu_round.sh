#!/bin/bash
# tests + bench + microbenchmarks on the GPU box
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" > gpurun_out/phases.txt
tail -3 gpurun_out/pytest_gpu.log
grep -E "^(FAILED|E  )|teacher-forced|flip rate" gpurun_out/pytest_gpu.log | cut -c1-250
for mb in 32 64; do
  for rep in a b; do
  timeout 600 python bench.py --steps 30 --warmup 5 --micro-batch $mb --no-cpu-baseline > gpurun_out/bench_mb${mb}_$rep.json 2>> gpurun_out/bench.err
  echo "bench mb$mb $rep exit $?" >> gpurun_out/phases.txt
  done
done
timeout 600 python bench.py --steps 30 --warmup 5 --micro-batch 64 --no-cpu-baseline --no-clocks > gpurun_out/bench_mb64_noclk.json 2>> gpurun_out/bench.err
timeout 600 python tools/bench_fusion.py 64 50 > gpurun_out/bench_fusion.txt 2>&1
timeout 600 python tools/bench_latency.py 300 > gpurun_out/bench_latency.txt 2>&1
cat gpurun_out/bench_fusion.txt gpurun_out/bench_latency.txt
cat gpurun_out/phases.txt; for f in gpurun_out/bench_mb*_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print('value %.0f e2e %.0f ms %.2f tc_frac %.3f share %.2f tc_ms %.2f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['roofline']['kernel_ms_per_step']), d['step_ms'], d['clocks'])
"; done
