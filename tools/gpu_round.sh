#!/bin/bash
# tests + bench + ncu launch list on the GPU box
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" > gpurun_out/phases.txt
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/phases.txt
cp gpurun_out/tc_launches.csv gpurun_out/tc_launches_mb16.csv 2>/dev/null
for mb in 4 8 32 64; do
  timeout 600 python bench.py --steps 10 --warmup 3 --micro-batch $mb --no-cpu-baseline > gpurun_out/bench_mb$mb.json 2>> gpurun_out/bench.err
  echo "bench mb$mb exit $?" >> gpurun_out/phases.txt
done
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 864 -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu exit $?" >> gpurun_out/phases.txt
cat gpurun_out/phases.txt; cat gpurun_out/bench.json
