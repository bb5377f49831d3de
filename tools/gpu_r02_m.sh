#!/bin/bash
# full GPU suite + smoke + default bench with the CTA-pair variants on by default; A/B against HMV_PAIR=0 HMV_SEAM_PAIR=0
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"
tail -2 gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_gpu.log | cut -c1-400
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
Q="--steps 30 --warmup 3 --no-e2e --no-eager --no-latency --no-cpu-baseline"
for e in "HMV_PAIR=0 HMV_SEAM_PAIR=0" "HMV_PAIR=1" "HMV_PAIR=0 HMV_SEAM_PAIR=0" "HMV_PAIR=1"; do
  env $e timeout 300 python bench.py $Q > gpurun_out/bench_v.json 2>/dev/null
  python - "$e" <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_v.json")); r = d["roofline"]
print("%-28s value %.0f step median %.3f mean %.3f" % (sys.argv[1], d["value"], d["step_ms"]["median"], d["ms_per_step"]))
PY
done
python tools/bench_latency.py 300 | head -2
