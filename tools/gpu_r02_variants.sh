#!/bin/bash
# variant libraries (tools/build_variants.py: -D knobs) selected with HMV_LIB_PATH: per-step parity test + per-class times
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
g = lambda k: cl[k]["ms_per_launch"] if k in cl else float("nan")
print("%-40s test rc %s | seam %.4f  l1 tail %.4f  l2 tail %.4f  l3.conv2 %.4f  l1.conv1 %.4f | step median %.3f" % (sys.argv[1][-38:], sys.argv[2], s["ms_per_launch"], g("layer1.x.conv2+conv3"), g("layer2.x.conv2+conv3"), g("layer3.x.conv2"), g("layer1.x.conv1"), d["step_ms"]["median"]))
PY
done
