#!/bin/bash
# tests + bench + conv microbench (+ optional ncu) on the GPU box
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" > gpurun_out/phases.txt
tail -3 gpurun_out/pytest_gpu.log
grep -E "^(FAILED|E  )|teacher-forced|flip rate|well-cond" gpurun_out/pytest_gpu.log | cut -c1-250
for mb in 32 64; do
  for rep in a b; do
  timeout 600 python bench.py --steps 20 --warmup 5 --micro-batch $mb --no-cpu-baseline > gpurun_out/bench_mb${mb}_$rep.json 2>> gpurun_out/bench.err
  echo "bench mb$mb $rep exit $?" >> gpurun_out/phases.txt
  done
  cp gpurun_out/tc_launches.csv gpurun_out/tc_launches_mb$mb.csv 2>/dev/null
done
timeout 600 python tools/bench_conv.py 320 > gpurun_out/bench_conv.txt 2>&1
cat gpurun_out/bench_conv.txt
if [ "$1" == "ncu" ]; then
for c in l3.conv3; do
timeout 300 python tools/bench_conv.py 160 $c 1 > gpurun_out/plain_$c.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tc -s 1 -c 1 -o gpurun_out/prof_$c -f \
    python tools/bench_conv.py 160 $c 1 > gpurun_out/ncu_$c.log 2>&1
echo "ncu $c exit $?" >> gpurun_out/phases.txt
done
fi
cat gpurun_out/phases.txt; for f in gpurun_out/bench_mb*_?.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print('value %.0f e2e %.0f ms %.2f tc_frac %.3f share %.2f tc_ms %.2f'%(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_share_of_step'], d['roofline']['kernel_ms_per_step']), d['clocks'])
"; done
