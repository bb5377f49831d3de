#!/bin/bash
# round 2: full GPU suite + default bench + HRNet bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"
tail -3 gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_gpu.log | cut -c1-400
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc $?"; tail -3 gpurun_out/bench_default.err
timeout 900 python bench.py --backbone hrnet --steps 20 --no-cpu-baseline > gpurun_out/bench_hrnet.json 2> gpurun_out/bench_hrnet.err; echo "bench hrnet rc $?"; tail -3 gpurun_out/bench_hrnet.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_default.json", "gpurun_out/bench_hrnet.json"):
    try: d = json.load(open(f))
    except Exception as e: print(f, "unreadable", e); continue
    r = d["roofline"]
    print(f, "value %.1f ms/step %.3f median %.3f max %.2f e2e %.1f fp32-in %.1f launches %d" % (d["value"], d["ms_per_step"], d["step_ms"]["median"], d["step_ms"]["max"], d["e2e"]["value"], d["e2e"]["fp32_input"]["value"], d["gpu_launches"]))
    print("  dominant", r["kernel"], r["bound"], "frac %.3f" % r["frac"], "traffic", r.get("traffic"), r.get("traffic_note"), "all_tc frac %.3f" % r["all_tc"]["frac"], "model TF %.0f" % r["all_tc"]["end_to_end_model_tflops"], "phases", {k: round(v, 3) for k, v in r["phase_ms_per_step"].items()})
    e = d.get("gpu_eager_baseline")
    if e: print("  eager fp32 %.1f ms bf16 %.1f ms -> x%.2f / x%.2f" % (e["fp32_tf32conv"]["ms_per_step"], e["bf16_autocast"]["ms_per_step"], e["speedup_vs_eager_fp32"], e["speedup_vs_eager_bf16"]))
    print("  latency", d.get("latency_b1"), "clocks", (d.get("clocks") or {}).get("sm_mhz"), (d.get("clocks") or {}).get("samples"))
PY
