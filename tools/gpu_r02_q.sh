#!/bin/bash
# slow-step hunt: the default bench loop with (a) no sampler, (b) the sampler, (c) NVML initialised only after the enqueue
mkdir -p gpurun_out
Q="--steps 20 --warmup 3 --no-e2e --no-eager --no-latency --no-cpu-baseline"
run() {
  env $1 HMV_PAIR_MINK=512 timeout 300 python bench.py $Q $2 > gpurun_out/bench_v.json 2>/dev/null
  python - "$1 $2" <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_v.json")); c = d.get("clocks") or {}
print("%-40s value %.0f step median %.3f max %.3f | enqueue %.2f ms/step | nvml samples %s" % (sys.argv[1], d["value"], d["step_ms"]["median"], d["step_ms"]["max"], d["step_ms"]["cpu_enqueue"], c.get("samples")))
PY
}
for i in 1 2 3 4 5; do run "A=1" "--no-clocks"; run "A=1" ""; run "HMV_BENCH_SAMPLER_LATE=1" ""; done
