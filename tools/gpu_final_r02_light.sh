#!/bin/bash
# end of round 2: full GPU suite, smoke, ncu launch list + per-class DRAM traffic + full captures of the final kernels,
# then the default / HRNet / reference-arm bench lines with the refreshed traffic file in place
mkdir -p gpurun_out
tools/ncu_capture_r02.sh > gpurun_out/ncu_capture.log 2>&1; grep -E "rc |tc_launches_per_step|launches_per_step" gpurun_out/ncu_capture.log | head -12
cp gpurun_out/r02/tc_traffic.json profiles/r02/tc_traffic.json
timeout 900 python bench.py > gpurun_out/bench_1gpu_default.json 2> gpurun_out/bench_default.err; echo "default rc $?"
timeout 900 python bench.py --backbone hrnet --no-cpu-baseline > gpurun_out/bench_1gpu_hrnet.json 2> gpurun_out/bench_hrnet.err; echo "hrnet rc $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_arm.json 2> gpurun_out/bench_ref.err; echo "reference rc $?"
python - <<'PY'
import json
for f in ("bench_1gpu_default", "bench_1gpu_hrnet", "bench_reference_arm"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
    except Exception as e:
        print(f, "unreadable", e); continue
    r = d.get("roofline") or {}
    print(f, "value %.1f ms/step %s e2e %s launches %s | roofline %s frac %s traffic %s" % (d["value"], d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"), r.get("kernel"), r.get("frac"), r.get("traffic")))
PY
