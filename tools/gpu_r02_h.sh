#!/bin/bash
# which kernels are chip-bandwidth bound and which are per-SM latency bound: the same step with 148 / 111 / 74 persistent CTAs
mkdir -p gpurun_out
Q="--steps 20 --warmup 3 --no-e2e --no-eager --no-latency --no-cpu-baseline --no-clocks"
for n in 148 111 74; do
  HMV_NUM_SMS=$n timeout 300 python bench.py $Q > gpurun_out/bench_h_$n.json 2>/dev/null
done
python - <<'PY'
import json
rows = {}
for n in (148, 111, 74):
    d = json.load(open(f"gpurun_out/bench_h_{n}.json"))
    for c in d["roofline"]["classes"]:
        rows.setdefault(c["kernel"], {})[n] = c["ms_per_launch"]
    rows.setdefault("STEP median", {})[n] = d["step_ms"]["median"]
print("%-30s %9s %9s %9s   t111/t148 t74/t148 (1.33 / 2.0 = per-SM bound, 1.0 = chip-bandwidth bound)" % ("kernel", 148, 111, 74))
for k, v in rows.items():
    print("%-30s %9.4f %9.4f %9.4f   %.2f %.2f" % (k, v[148], v[111], v[74], v[111] / v[148], v[74] / v[148]))
PY
