#!/bin/bash
# end-of-round validation: GPU tests, smoke, default bench (both arms) + same-box A/B of the env switches given in AB_ENVS
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"
tail -2 gpurun_out/pytest_gpu.log
grep -E "^(FAILED|E  )|teacher-forced|flip rate|eager-on-GPU|preprocess kernel" gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -4
timeout 600 python bench.py --impl reference --steps 10 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref rc $?"
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc $?"
cp gpurun_out/tc_launches.csv gpurun_out/tc_launches_default.csv
for e in $AB_ENVS; do
  env $e timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_ab_$e.json 2>/dev/null; echo "ab $e rc $?"
done
python - <<'PY'
import json, glob
for f in ["gpurun_out/bench_default.json", "gpurun_out/bench_ref.json"] + sorted(glob.glob("gpurun_out/bench_ab_*.json")):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, "unreadable", e); continue
    e2e = d.get("e2e") or {}
    print(f, "value %.1f %s ms/step %.2f median %s" % (d["value"], d["unit"], d["ms_per_step"], (d.get("step_ms") or {}).get("median")), "e2e", e2e.get("value"), "u8", (e2e.get("uint8_input") or {}).get("value"),
          "frac", (d.get("roofline") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "launches", d.get("gpu_launches"))
PY
