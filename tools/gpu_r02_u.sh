#!/bin/bash
# fusion layer: one fused kernel (mma.sync) vs the un-fused form whose GEMMs run on the tcgen05 kernel (lean issue path now)
mkdir -p gpurun_out
Q="--steps 30 --warmup 3 --no-e2e --no-eager --no-latency --no-cpu-baseline"
for e in "A=1" "HMV_FUSION_UNFUSED=1" "A=1" "HMV_FUSION_UNFUSED=1"; do
  env $e timeout 300 python bench.py $Q > gpurun_out/bench_v.json 2>/dev/null
  python - "$e" <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_v.json")); r = d["roofline"]
print("%-24s value %.0f step median %.3f | phases %s | launches %s" % (sys.argv[1], d["value"], d["step_ms"]["median"], {k: round(v, 3) for k, v in r["phase_ms_per_step"].items()}, d["gpu_launches"]))
PY
done
python tools/bench_latency.py 300 | head -1 | cut -c1-200
HMV_FUSION_UNFUSED=1 python tools/bench_latency.py 300 | head -1 | cut -c1-200
