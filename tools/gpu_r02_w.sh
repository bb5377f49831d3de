#!/bin/bash
# tail kernel CTA-pair variant: full GPU suite, smoke, A/B against HMV_TAIL_PAIR=0
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"
tail -1 gpurun_out/pytest_gpu.log; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.log | head
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
Q="--steps 30 --warmup 3 --no-e2e --no-eager --no-latency --no-cpu-baseline"
for e in "HMV_TAIL_PAIR=1" "HMV_TAIL_PAIR=0" "HMV_TAIL_PAIR=1" "HMV_TAIL_PAIR=0"; do
  env $e timeout 300 python bench.py $Q > gpurun_out/bench_v.json 2>/dev/null
  python - "$e" <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_v.json")); r = d["roofline"]
cl = {c["kernel"]: c for c in r["classes"]}
print("%-18s value %.0f step median %.3f | " % (sys.argv[1], d["value"], d["step_ms"]["median"]) + "  ".join("%s %.4f" % (k.replace("layer", "l").replace(".x.", "."), v["ms_per_launch"]) for k, v in cl.items() if "conv2+conv3" in k))
PY
done
