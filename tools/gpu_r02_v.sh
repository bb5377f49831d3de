#!/bin/bash
# layer1.0: downsample folded into the fused tail's conv3 (HMV_FOLD_DS=0 switches it off): parity + A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -x -k "not hrnet" > gpurun_out/pytest_s.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/pytest_s.log | cut -c1-300
Q="--steps 30 --warmup 3 --no-e2e --no-eager --no-latency --no-cpu-baseline"
for e in "HMV_FOLD_DS=1" "HMV_FOLD_DS=0" "HMV_FOLD_DS=1" "HMV_FOLD_DS=0"; do
  env $e timeout 300 python bench.py $Q > gpurun_out/bench_v.json 2>gpurun_out/bench_v.err || tail -3 gpurun_out/bench_v.err
  python - "$e" <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_v.json")); r = d["roofline"]
cl = {c["kernel"]: c for c in r["classes"]}
print("%-16s value %.0f step median %.3f | " % (sys.argv[1], d["value"], d["step_ms"]["median"]) + "  ".join("%s %.4f" % (k.replace("layer", "l").replace(".x.", "."), v["ms_per_launch"]) for k, v in cl.items() if k.startswith("layer1")) + " | launches %d" % d["gpu_launches"])
PY
done
