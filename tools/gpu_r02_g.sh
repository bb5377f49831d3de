#!/bin/bash
# round 2, call G: biases through the constant bank - correctness + timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -rA -p no:cacheprovider -k "conv_bn_act or steps or bench_configuration or stagewise or fusion_stage" > gpurun_out/pytest_gpu_g.log 2>&1; echo "pytest rc $?"
tail -3 gpurun_out/pytest_gpu_g.log
grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_gpu_g.log | cut -c1-600
P="--steps 3 --warmup 3 --ramp-seconds 0.5 --no-e2e --no-eager --no-latency --no-cpu-baseline --no-clocks"
HMV_BN_PROF=1 timeout 300 python bench.py $P 2> gpurun_out/bn_prof_g.err > /dev/null; grep bn_prof gpurun_out/bn_prof_g.err | head -2
HMV_BT_PROF=1 timeout 300 python bench.py $P 2> gpurun_out/bt_prof_g.err > /dev/null; grep bt_prof gpurun_out/bt_prof_g.err | sed -n '1p;5p'
Q="--steps 30 --warmup 3 --no-e2e --no-eager --no-latency --no-cpu-baseline"
for i in 1 2; do
  timeout 300 python bench.py $Q > gpurun_out/bench_g_$i.json 2>/dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_g_*.json")):
    d = json.load(open(f)); r = d["roofline"]
    print(f, "value %.0f" % d["value"], "median %.3f max %.2f" % (d["step_ms"]["median"], d["step_ms"]["max"]), "phases", {k: round(v, 3) for k, v in r["phase_ms_per_step"].items()})
    for c in r["classes"]: print("   %-28s x%-2d %.4f ms  %s frac %.3f" % (c["kernel"], c["launches_per_step"], c["ms_per_launch"], c["bound"], c["frac"]))
PY
