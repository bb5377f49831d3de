#!/bin/bash
# round 2, call C: tests, seam stall counters, tail stall counters
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"
tail -3 gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_gpu.log | cut -c1-400
HMV_BN_PROF=1 timeout 300 python bench.py --steps 3 --warmup 3 --ramp-seconds 0.5 --no-e2e --no-eager --no-latency --no-cpu-baseline --no-clocks 2> gpurun_out/bn_prof.err > /dev/null; grep bn_prof gpurun_out/bn_prof.err | head -6
HMV_BT_PROF=1 timeout 300 python bench.py --steps 3 --warmup 3 --ramp-seconds 0.5 --no-e2e --no-eager --no-latency --no-cpu-baseline --no-clocks 2> gpurun_out/bt_prof.err > /dev/null; grep bt_prof gpurun_out/bt_prof.err | head -8
for i in 1 2 3; do
  timeout 300 python bench.py --steps 40 --no-e2e --no-eager --no-latency --no-cpu-baseline > gpurun_out/bench_gc_$i.json 2>/dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_gc_*.json")):
    d = json.load(open(f)); c = d.get("clocks") or {}
    print(f, "value %.0f" % d["value"], d["step_ms"])
PY
