#!/bin/bash
# late round 2: small-pass shapes (graph head 16 x 16 / 64 x 8 K-splits, N = 128 tiles for the stand-alone layer3 convs of a
# B = 1 pass): full GPU suite + smoke, B = 1 latency A/B against HMV_GCN_SMALL=0 HMV_NARROW_SMALL=0, then the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"
tail -1 gpurun_out/pytest_gpu.log; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.log | head
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
for e in "HMV_GCN_SMALL=1" "HMV_GCN_SMALL=0 HMV_NARROW_SMALL=0" "HMV_GCN_SMALL=0" "HMV_NARROW_SMALL=0"; do
  echo "== $e"; env $e timeout 200 python tools/bench_latency.py 200 2>&1 | tail -3
done | tee gpurun_out/latency_ab.txt
timeout 600 python bench.py > gpurun_out/bench_1gpu_default.json 2> gpurun_out/bench_default.err; echo "default rc $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_1gpu_default.json"))
r = d.get("roofline") or {}
print("value %.1f ms/step %s e2e %s launches %s | roofline %s frac %s traffic %s | latency %s | eager %s" % (d["value"], d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"), r.get("kernel"), r.get("frac"), r.get("traffic"), {k: v.get("p50_ms") for k, v in (d.get("latency_b1") or {}).items() if isinstance(v, dict)}, {k: v for k, v in (d.get("gpu_eager_baseline") or {}).items() if k.startswith("speedup")}))
PY
