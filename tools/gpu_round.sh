#!/bin/bash
# tests + bench + ncu launch list on the GPU box
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" > gpurun_out/phases.txt
tail -3 gpurun_out/pytest_gpu.log
for mb in 16 32 64; do
  timeout 600 python bench.py --steps 10 --warmup 3 --micro-batch $mb --no-cpu-baseline > gpurun_out/bench_mb$mb.json 2>> gpurun_out/bench.err
  echo "bench mb$mb exit $?" >> gpurun_out/phases.txt
  cp gpurun_out/tc_launches.csv gpurun_out/tc_launches_mb$mb.csv 2>/dev/null
done
if [ "$1" == "ncu" ]; then
timeout 600 python bench.py --steps 2 --warmup 3 --micro-batch 32 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 450 -c 160 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --micro-batch 32 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/phases.txt
fi
cat gpurun_out/phases.txt; cat gpurun_out/bench_mb32.json
