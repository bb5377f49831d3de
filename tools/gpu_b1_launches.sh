#!/bin/bash
# per-kernel durations of one B = 1 forward (graphs off, ncu --metrics gpu__time_duration.sum --clock-control none)
O=gpurun_out/r02; mkdir -p $O
HMV_NO_GRAPH=1 python tools/b1_forward.py 5 6 > $O/plain_b1.log 2>&1 || { echo "plain run failed"; tail -5 $O/plain_b1.log; exit 1; }
HMV_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/ncu_launches_b1.csv python tools/b1_forward.py 5 6 > $O/ncu_b1.log 2>&1; echo "ncu b1 rc $?"
python - <<'PY' | tee gpurun_out/r02/ncu_launches_b1_summary.txt
import csv, collections
rows = [ln for ln in open("gpurun_out/r02/ncu_launches_b1.csv") if ln.startswith('"')]
ks = [r for r in csv.DictReader(rows) if r["Metric Name"] == "gpu__time_duration.sum" and "hmv::" in r["Kernel Name"]]
per = len(ks) // 6
last = ks[-per:]
tot = 0.0
agg = collections.OrderedDict()
seq = []
for r in last:
    us = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1.0)
    name = r["Kernel Name"].split("(")[0].replace("void hmv::", "").replace("<unnamed>::", "")[:60]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us; tot += us
    seq.append((name, us, r.get("Grid Size", ""), r.get("Block Size", "")))
print("B=1 (5 views) serialised kernel time %.1f us over %d kernels" % (tot, len(last)))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]): print("  %-62s x%-2d %8.1f us" % (k, v[0], v[1]))
print("in launch order:")
for i, (n, us, g, b) in enumerate(seq): print("  %2d %-58s %7.1f us  grid %s block %s" % (i, n, us, g, b))
PY
