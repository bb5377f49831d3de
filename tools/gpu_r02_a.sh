#!/bin/bash
# round 2, call A: GPU tests, smoke, both bench arms, ncu launch list of one step
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 1500 python -m pytest tests -q -m gpu --no-header -rA -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"
tail -3 gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR|E  )|teacher-forced|chained|graph replay|step outputs|flip rate|eager-on-GPU|preprocess kernel" gpurun_out/pytest_gpu.log | cut -c1-1500
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -4
timeout 600 python bench.py --impl reference --steps 10 > gpurun_out/bench_ref.json 2>/dev/null; echo "ref rc $?"
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc $?"; tail -5 gpurun_out/bench_default.err
cp gpurun_out/tc_launches.csv gpurun_out/tc_launches_default.csv
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_default.json"))
r = d["roofline"]
print("value %.1f ms/step %.3f median %.3f e2e %.1f fp32-in %.1f" % (d["value"], d["ms_per_step"], d["step_ms"]["median"], d["e2e"]["value"], d["e2e"]["fp32_input"]["value"]))
print("dominant", r["kernel"], r["bound"], "frac %.3f" % r["frac"], "all_tc frac %.3f" % r["all_tc"]["frac"], "phases", r["phase_ms_per_step"])
for c in r["classes"]: print("  ", c)
print("eager", d.get("gpu_eager_baseline"))
print("latency", d.get("latency_b1"))
print("cpu", d.get("cpu_baseline"))
print("clocks", d.get("clocks"))
PY
B="python bench.py --steps 1 --warmup 3 --ramp-seconds 0 --no-e2e --no-eager --no-latency --no-cpu-baseline --no-clocks"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/ncu_step.csv $B > gpurun_out/ncu_step.log 2>&1; echo "ncu rc $?"
python tools/ncu_step_traffic.py gpurun_out/ncu_step.csv gpurun_out/tc_launches.csv gpurun_out/tc_traffic.json 5
