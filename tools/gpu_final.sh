#!/bin/bash
# end-of-round validation: GPU tests, smoke, default bench (both arms), 8-view and large-batch lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"
tail -2 gpurun_out/pytest_gpu.log
grep -E "^(FAILED|E  )|teacher-forced|flip rate|eager-on-GPU|preprocess kernel" gpurun_out/pytest_gpu.log | cut -c1-300
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -4
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2>/dev/null; echo "ref rc $?"
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc $?"
cp gpurun_out/tc_launches.csv gpurun_out/tc_launches_default.csv
timeout 600 python bench.py --views 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_views8.json 2>/dev/null; echo "views8 rc $?"
timeout 600 python bench.py --batch 1024 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_b1024.json 2>/dev/null; echo "b1024 rc $?"
python - <<'PY'
import json
for f in ("bench_default", "bench_views8", "bench_b1024", "bench_ref"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
    except Exception as e:
        print(f, "unreadable", e); continue
    e2e = d.get("e2e") or {}
    print(f, "value %.1f %s ms/step %.2f" % (d["value"], d["unit"], d["ms_per_step"]), "e2e", e2e.get("value"), "u8", (e2e.get("uint8_input") or {}).get("value"),
          "frac", (d.get("roofline") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "launches", d.get("gpu_launches"))
PY
