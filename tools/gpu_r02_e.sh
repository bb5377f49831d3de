#!/bin/bash
# round 2, call E: seam kernel with the residual read by the epilogue warps (A/B against the TMA slot prefetch)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -rA -p no:cacheprovider -k "steps or bench_configuration" > gpurun_out/pytest_gpu_e.log 2>&1; echo "pytest rc $?"
tail -3 gpurun_out/pytest_gpu_e.log
grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_gpu_e.log | cut -c1-600
P="--steps 3 --warmup 3 --ramp-seconds 0.5 --no-e2e --no-eager --no-latency --no-cpu-baseline --no-clocks"
HMV_BN_PROF=1 timeout 300 python bench.py $P 2> gpurun_out/bn_prof_direct.err > /dev/null; grep bn_prof gpurun_out/bn_prof_direct.err | head -2
HMV_BN_RES_DIRECT=0 HMV_BN_PROF=1 timeout 300 python bench.py $P 2> gpurun_out/bn_prof_tma.err > /dev/null; grep bn_prof gpurun_out/bn_prof_tma.err | head -2
Q="--steps 30 --warmup 3 --no-e2e --no-eager --no-latency --no-cpu-baseline"
for i in 1 2; do
  timeout 300 python bench.py $Q > gpurun_out/bench_e_direct1_$i.json 2>/dev/null
  HMV_BN_RES_DIRECT=0 timeout 300 python bench.py $Q > gpurun_out/bench_e_direct0_$i.json 2>/dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_e_*.json")):
    d = json.load(open(f)); r = d["roofline"]
    cl = {c["kernel"]: c for c in r["classes"]}
    print(f, "value %.0f" % d["value"], "median %.3f max %.2f" % (d["step_ms"]["median"], d["step_ms"]["max"]),
          " ".join("%s %.4f" % (k.split(".x.")[-1][:14] + "@" + k[:6], cl[k]["ms_per_launch"]) for k in ("layer3.x.conv3+next.conv1", "layer1.x.conv2+conv3", "layer2.x.conv2+conv3", "layer3.x.conv2") if k in cl))
PY
