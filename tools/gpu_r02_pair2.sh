#!/bin/bash
# cta_group::2 pair variant of the layer3 seam kernel: correctness (per-step teacher-forced, B=64 replay), then A/B
mkdir -p gpurun_out
export HMV_SEAM_PAIR=1
timeout 600 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -x -k "steps_teacher_forced and bf16 and not hrnet" > gpurun_out/pair_seam_test.log 2>&1; echo "pair seam step tests rc $?"; tail -2 gpurun_out/pair_seam_test.log; grep -E "^E  " gpurun_out/pair_seam_test.log | head -6 | cut -c1-400
timeout 600 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -x -k "bench_configuration or stagewise_teacher_forced" > gpurun_out/pair_seam_test2.log 2>&1; echo "pair seam model tests rc $?"; tail -2 gpurun_out/pair_seam_test2.log; grep -E "^E  " gpurun_out/pair_seam_test2.log | head -6 | cut -c1-400
unset HMV_SEAM_PAIR
Q="--steps 30 --warmup 3 --no-e2e --no-eager --no-latency --no-cpu-baseline"
for e in HMV_SEAM_PAIR=0 HMV_SEAM_PAIR=1 HMV_SEAM_PAIR=0 HMV_SEAM_PAIR=1; do
  env $e timeout 300 python bench.py $Q > gpurun_out/bench_v.json 2>/dev/null
  python - "$e" <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_v.json")); r = d["roofline"]
cl = {c["kernel"]: c for c in r["classes"]}
s = cl["layer3.x.conv3+next.conv1"]
print("%-18s step median %.3f | seam %.4f ms issue frac %.3f" % (sys.argv[1], d["step_ms"]["median"], s["ms_per_launch"], s.get("frac_issue", 0)))
PY
done
HMV_SEAM_PAIR=1 HMV_BN_PROF=1 timeout 300 python bench.py --steps 3 --warmup 3 --ramp-seconds 0.5 --no-e2e --no-eager --no-latency --no-cpu-baseline --no-clocks 2> gpurun_out/bn_prof_pair.err > /dev/null; grep bn_prof gpurun_out/bn_prof_pair.err | head -2
