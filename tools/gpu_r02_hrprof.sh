#!/bin/bash
# where the HRNet step goes: ncu launch list (durations) of whole B=64 steps, aggregated per kernel
mkdir -p gpurun_out/r02
B="python bench.py --backbone hrnet --steps 1 --warmup 3 --ramp-seconds 0 --no-e2e --no-eager --no-latency --no-cpu-baseline --no-clocks"
$B > gpurun_out/r02/plain_hr.log 2>&1 || { echo plain failed; tail -5 gpurun_out/r02/plain_hr.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02/ncu_launches_hrnet.csv $B > gpurun_out/r02/ncu_hr.log 2>&1; echo "ncu rc $?"
python - <<'PY'
import csv, collections, re
rows = [ln for ln in open("gpurun_out/r02/ncu_launches_hrnet.csv") if ln.startswith('"')]
ks = [r for r in csv.DictReader(rows) if "hmv::" in r["Kernel Name"]]
per = len(ks) // 5
step = ks[-per:]
agg = collections.OrderedDict(); tot = 0.0
for r in step:
    us = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1.0)
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void hmv::", "").replace("<unnamed>::", "")[:70]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us; tot += us
print("HRNet B=64 step: %d kernels, %.2f ms serialised" % (per, tot * 1e-3))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]): print("  %-72s x%-3d %9.1f us" % (k, v[0], v[1]))
PY
