#!/bin/bash
# small-pass graph head (all weight loads up front) + short stem strips: full GPU suite, smoke, B = 1 latency A/B, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"
tail -1 gpurun_out/pytest_gpu.log; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.log | head
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
for e in "HMV_GCN_SMALL=1" "HMV_GCN_SMALL=0"; do
  echo "== $e"; env $e timeout 200 python tools/bench_latency.py 200 2>&1 | tail -3
done | tee gpurun_out/latency_gcn_ab.txt
tools/gpu_b1_launches.sh 2>&1 | head -24
