#!/bin/bash
# seam kernel: operand stages / staging slots / conv1 lag sweep (variant libraries built with -DHMV_BN_*), selected with HMV_LIB_PATH
mkdir -p gpurun_out
Q="--steps 30 --warmup 3 --no-e2e --no-eager --no-latency --no-cpu-baseline"
for lib in default handmvnet_b200/lib/variants/*.so default; do
  if [ "$lib" = default ]; then unset HMV_LIB_PATH; else export HMV_LIB_PATH=$PWD/$lib; fi
  timeout 300 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -k "steps_teacher_forced and bf16-default and not hrnet" > gpurun_out/v_test.log 2>&1; rc=$?
  timeout 300 python bench.py $Q > gpurun_out/bench_v.json 2>/dev/null
  python - "$lib" "$rc" <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_v.json")); r = d["roofline"]
cl = {c["kernel"]: c for c in r["classes"]}
s = cl["layer3.x.conv3+next.conv1"]
print("%-52s test rc %s | seam %.4f ms (issue frac %.3f) | step median %.3f" % (sys.argv[1][-46:], sys.argv[2], s["ms_per_launch"], s.get("frac_issue", 0), d["step_ms"]["median"]))
PY
done
