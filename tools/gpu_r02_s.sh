#!/bin/bash
# seam kernel with the store warp: parity (conv cases, per-step, B=64 replay, stagewise), bench A/B, stall counters
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -x -k "not hrnet" > gpurun_out/pytest_s.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_s.log
Q="--steps 30 --warmup 3 --no-e2e --no-eager --no-latency --no-cpu-baseline"
for i in 1 2; do
  timeout 300 python bench.py $Q > gpurun_out/bench_v.json 2>/dev/null
  python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_v.json")); r = d["roofline"]
cl = {c["kernel"]: c for c in r["classes"]}
g = lambda k: cl[k]["ms_per_launch"] if k in cl else float("nan")
print("value %.0f step median %.3f max %.3f | seam %.4f  l1 tail %.4f  l2 tail %.4f  l3.conv2 %.4f" % (d["value"], d["step_ms"]["median"], d["step_ms"]["max"], g("layer3.x.conv3+next.conv1"), g("layer1.x.conv2+conv3"), g("layer2.x.conv2+conv3"), g("layer3.x.conv2")))
PY
done
HMV_BT_PROF=1 timeout 300 python bench.py --steps 3 --warmup 3 --ramp-seconds 0.5 --no-e2e --no-eager --no-latency --no-cpu-baseline --no-clocks 2> gpurun_out/bt_prof_s.err > /dev/null; grep bt_prof gpurun_out/bt_prof_s.err | head -5 | cut -c1-600
