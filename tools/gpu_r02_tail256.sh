#!/bin/bash
# layer3 as conv2+conv3 tails (P = 256) + separate conv1 launches instead of conv2 launches + conv3/conv1 seams
mkdir -p gpurun_out
Q="--steps 30 --warmup 3 --no-e2e --no-eager --no-latency --no-cpu-baseline"
for e in HMV_FUSE_TAIL=3 HMV_FUSE_TAIL=7; do
  env $e timeout 300 python bench.py $Q > gpurun_out/bench_t_$e.json 2>/dev/null
  python - "$e" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_t_{sys.argv[1]}.json")); r = d["roofline"]
print(sys.argv[1], "value %.0f median %.3f" % (d["value"], d["step_ms"]["median"]), "backbone %.3f" % r["phase_ms_per_step"]["backbone"])
for c in r["classes"]:
    if c["kernel"].startswith("layer3"): print("   %-28s x%-2d %.4f ms  %s frac %.3f  %.0f TF" % (c["kernel"], c["launches_per_step"], c["ms_per_launch"], c["bound"], c["frac"], c["tflops"]))
PY
done
