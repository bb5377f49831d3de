#!/bin/bash
# fusion layer as clusters of 8 CTAs for small passes: the tests that reach it + B = 1 latency A/B against HMV_FUSION_CLUSTER=0
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --no-header -rA -s -p no:cacheprovider -k "fusion or stagewise or chained or known or micro_batching or uint8 or golden_fixtures or end_to_end" > gpurun_out/pytest_fusion.log 2>&1; echo "pytest rc $?"
tail -1 gpurun_out/pytest_fusion.log; grep -E "^(FAILED|ERROR)|cluster kernel vs|teacher-forced stage errors" gpurun_out/pytest_fusion.log | head -30
for e in "HMV_FUSION_CLUSTER=1" "HMV_FUSION_CLUSTER=0" "HMV_FUSION_CLUSTER=1"; do
  echo "== $e"; env $e timeout 200 python tools/bench_latency.py 200 2>&1 | tail -3
done | tee gpurun_out/latency_fusion_ab.txt
