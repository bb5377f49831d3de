#!/bin/bash
# end of round 2 (second session): full GPU suite, smoke, the default and HRNet bench lines of the final build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"
tail -1 gpurun_out/pytest_gpu.log; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_gpu.log | head
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_1gpu_default.json 2> gpurun_out/bench_default.err; echo "default rc $?"
timeout 600 python bench.py --backbone hrnet --no-cpu-baseline > gpurun_out/bench_1gpu_hrnet.json 2> gpurun_out/bench_hrnet.err; echo "hrnet rc $?"
python - <<'PY'
import json
for f in ("bench_1gpu_default", "bench_1gpu_hrnet"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
    except Exception as e:
        print(f, "unreadable", e); continue
    r = d.get("roofline") or {}
    print(f, "value %.1f ms/step %s e2e %s launches %s | roofline %s frac %s traffic %s | latency %s | eager %s" % (d["value"], d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), d.get("gpu_launches"), r.get("kernel"), r.get("frac"), r.get("traffic"), {k: v.get("p50_ms") for k, v in (d.get("latency_b1") or {}).items() if isinstance(v, dict)}, {k: v for k, v in (d.get("gpu_eager_baseline") or {}).items() if k.startswith("speedup")}))
PY
