#!/bin/bash
# lean MMA issue path (whole warp + elect.sync): full GPU suite, smoke, bench, stall counters
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"
tail -2 gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_gpu.log | cut -c1-400 | head -20
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
Q="--steps 30 --warmup 3 --no-e2e --no-eager --no-latency --no-cpu-baseline"
for e in "HMV_PAIR=1" "HMV_PAIR=0 HMV_SEAM_PAIR=0" "HMV_PAIR=1"; do
  env $e timeout 300 python bench.py $Q > gpurun_out/bench_v.json 2>/dev/null
  python - "$e" <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_v.json")); r = d["roofline"]
print("%-28s value %.0f step median %.3f mean %.3f" % (sys.argv[1], d["value"], d["step_ms"]["median"], d["ms_per_step"]))
for c in r["classes"][:12]:
    print("   %-28s x%d %.4f ms  hbm/tensor frac %.3f" % (c["kernel"], c["launches_per_step"], c["ms_per_launch"], c["frac"]))
PY
done
cp gpurun_out/bench_v.json gpurun_out/bench_lean.json
HMV_BT_PROF=1 timeout 300 python bench.py --steps 3 --warmup 3 --ramp-seconds 0.5 --no-e2e --no-eager --no-latency --no-cpu-baseline --no-clocks 2> gpurun_out/bt_prof_o.err > /dev/null; grep bt_prof gpurun_out/bt_prof_o.err | head -5 | cut -c1-500
HMV_BN_PROF=1 timeout 300 python bench.py --steps 3 --warmup 3 --ramp-seconds 0.5 --no-e2e --no-eager --no-latency --no-cpu-baseline --no-clocks 2> gpurun_out/bn_prof_o.err > /dev/null; grep bn_prof gpurun_out/bn_prof_o.err | head -2 | cut -c1-700
