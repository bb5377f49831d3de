#!/bin/bash
# refresh the headline lines after the lean-issue / resident-Y2 changes: default bench (all sections), HRNet, B=1 latency
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_1gpu_default.json 2> gpurun_out/bench_default.err; echo "default rc $?"
timeout 900 python bench.py --backbone hrnet > gpurun_out/bench_1gpu_hrnet.json 2> gpurun_out/bench_hrnet.err; echo "hrnet rc $?"
python tools/bench_latency.py 300 > gpurun_out/bench_latency.txt 2>&1; cat gpurun_out/bench_latency.txt | cut -c1-300
HMV_SEAM_PAIR=0 python tools/bench_latency.py 300 2>&1 | head -1 | cut -c1-300
python - <<'PY'
import json
for f in ("bench_1gpu_default", "bench_1gpu_hrnet"):
    d = json.load(open("gpurun_out/%s.json" % f))
    print(f, "value %.0f ms/step %.3f e2e %s eager %s latency %s cpu %s launches %s" % (d["value"], d["ms_per_step"], d.get("e2e", {}).get("value"), {k: v for k, v in (d.get("gpu_eager_baseline") or {}).items() if "ms" in k or "speed" in k}, (d.get("latency_b1") or {}), (d.get("cpu_baseline") or {}).get("value"), d.get("gpu_launches")))
PY
