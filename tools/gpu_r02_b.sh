#!/bin/bash
# round 2, call B: full GPU tests, sampler A/B, traffic json, compute-sanitizer
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"
tail -3 gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR|E  )|teacher-forced|chained|graph replay|step outputs|flip rate|preprocess kernel" gpurun_out/pytest_gpu.log | cut -c1-3000
for i in 1 2 3; do
  timeout 300 python bench.py --steps 40 --no-e2e --no-eager --no-latency --no-cpu-baseline > gpurun_out/bench_clk_$i.json 2>/dev/null
  timeout 300 python bench.py --steps 40 --no-e2e --no-eager --no-latency --no-cpu-baseline --no-clocks > gpurun_out/bench_noclk_$i.json 2>/dev/null
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*clk_*.json")):
    d = json.load(open(f)); c = d.get("clocks") or {}
    print(f, "value %.0f" % d["value"], d["step_ms"], "q_max_region", c.get("query_ms_max_in_region"), "samples", c.get("samples"))
PY
python tools/ncu_step_traffic.py gpurun_out/ncu_step.csv gpurun_out/tc_launches.csv gpurun_out/tc_traffic.json 5
S=/usr/local/cuda/bin/compute-sanitizer
for prec in bf16 fp32; do
  timeout 900 $S --tool memcheck --print-limit 20 python tools/sanitize_forward.py $prec 2 > gpurun_out/sanitize_memcheck_$prec.log 2>&1; echo "memcheck $prec rc $?"; tail -4 gpurun_out/sanitize_memcheck_$prec.log
done
timeout 900 $S --tool racecheck --print-limit 20 python tools/sanitize_forward.py bf16 2 > gpurun_out/sanitize_racecheck_bf16.log 2>&1; echo "racecheck rc $?"; tail -6 gpurun_out/sanitize_racecheck_bf16.log
timeout 900 $S --tool synccheck --print-limit 20 python tools/sanitize_forward.py bf16 2 > gpurun_out/sanitize_synccheck_bf16.log 2>&1; echo "synccheck rc $?"; tail -4 gpurun_out/sanitize_synccheck_bf16.log
