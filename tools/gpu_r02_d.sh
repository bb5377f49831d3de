#!/bin/bash
# round 2, call D: fused-kernel tests, stall counters, A/B of the L2 prefetch
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -rA -p no:cacheprovider -k "steps or bench_configuration or stagewise or micro_batching" > gpurun_out/pytest_gpu_d.log 2>&1; echo "pytest rc $?"
tail -3 gpurun_out/pytest_gpu_d.log
grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_gpu_d.log | cut -c1-600
Q="--steps 30 --warmup 3 --no-e2e --no-eager --no-latency --no-cpu-baseline"
HMV_BN_PROF=1 timeout 300 python bench.py --steps 3 --warmup 3 --ramp-seconds 0.5 --no-e2e --no-eager --no-latency --no-cpu-baseline --no-clocks 2> gpurun_out/bn_prof.err > /dev/null; grep bn_prof gpurun_out/bn_prof.err | head -3
HMV_BT_PROF=1 timeout 300 python bench.py --steps 3 --warmup 3 --ramp-seconds 0.5 --no-e2e --no-eager --no-latency --no-cpu-baseline --no-clocks 2> gpurun_out/bt_prof.err > /dev/null; grep bt_prof gpurun_out/bt_prof.err | sed -n '1p;4p;5p'
for i in 1 2; do
  timeout 300 python bench.py $Q > gpurun_out/bench_d_pf1_$i.json 2>/dev/null
  HMV_BN_PREFETCH=0 timeout 300 python bench.py $Q > gpurun_out/bench_d_pf0_$i.json 2>/dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_d_*.json")):
    d = json.load(open(f)); r = d["roofline"]
    cl = {c["kernel"]: c for c in r["classes"]}
    print(f, "value %.0f" % d["value"], "median %.3f max %.2f" % (d["step_ms"]["median"], d["step_ms"]["max"]),
          " ".join("%s %.4f" % (k.split(".x.")[-1][:14] + "@" + k[:6], cl[k]["ms_per_launch"]) for k in ("layer3.x.conv3+next.conv1", "layer1.x.conv2+conv3", "layer2.x.conv2+conv3", "layer3.x.conv2") if k in cl))
PY
