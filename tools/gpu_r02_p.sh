#!/bin/bash
# after the lean issue path: which layers gain from the CTA-pair conv GEMM now (HMV_PAIR_MINK sweep), per-layer table
mkdir -p gpurun_out
Q="--steps 20 --warmup 3 --no-e2e --no-eager --no-latency --no-cpu-baseline"
for mk in 1024 512 256 64; do
  HMV_PAIR_MINK=$mk timeout 300 python bench.py $Q > gpurun_out/bench_v.json 2>/dev/null
  python - "$mk" <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_v.json")); r = d["roofline"]
print("HMV_PAIR_MINK=%-5s value %.0f step median %.3f mean %.3f" % (sys.argv[1], d["value"], d["step_ms"]["median"], d["ms_per_step"]))
print("   " + "  ".join("%s %.4f" % (c["kernel"].replace("layer", "l").replace(".x.", "."), c["ms_per_launch"]) for c in r["classes"] if "+" not in c["kernel"]))
PY
done
