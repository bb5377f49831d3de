#!/bin/bash
# round 2 profile set (run on the GPU box, one GPU): every command is first run WITHOUT ncu (exit code checked)
#   1. launch list of whole B=64 steps with DRAM bytes -> gpurun_out/r02/ncu_launches_step_dram.csv + tc_traffic.json
#   2. ncu --set full of one launch of the dominant kernel classes -> gpurun_out/r02/*.ncu-rep + key-metric summaries
#   3. launch list (durations) of one B=1 forward
O=gpurun_out/r02; mkdir -p $O
B="python bench.py --steps 1 --warmup 3 --ramp-seconds 0 --no-e2e --no-eager --no-latency --no-cpu-baseline --no-clocks"
$B > $O/plain_bench.log 2>&1 || { echo "plain bench failed"; tail -5 $O/plain_bench.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/ncu_launches_step_dram.csv $B > $O/ncu_step.log 2>&1; echo "ncu launch list rc $?"
# 4 ramp calls + 3 warm-up + 1 timed + 1 per-launch-event pass = 9 identical steps in the capture
python tools/ncu_step_traffic.py $O/ncu_launches_step_dram.csv gpurun_out/tc_launches.csv $O/tc_traffic.json 9 | tail -22
cap() {  # name, kernel regex, launches to skip
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 -f -o $O/$1 $B > $O/$1.log 2>&1; echo "ncu $1 rc $?"
}
cap seam_l3 bottleneck_next_kernel 16
cap tail_p64 "bottleneck_tail_kernel<\(int\)64" 10
cap tail_p128 "bottleneck_tail_kernel<\(int\)128" 13
cap conv_l3_conv2 "conv_gemm_tc_kernel<\(int\)256, \(int\)1, \(int\)4>" 13
cap stem_pool stem_pool_kernel 3
cap fusion_block fusion_block_kernel 15
python tools/ncu_summary.py $O $O > $O/ncu_summary.log 2>&1; tail -12 $O/ncu_summary.log
for n in seam_l3 tail_p64; do ncu -i $O/$n.ncu-rep --page details > $O/ncu_details_$n.txt 2>/dev/null; done
L="python tools/bench_latency.py 20"
HMV_NO_GRAPH=1 $L > $O/plain_latency.log 2>&1 && HMV_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/ncu_launches_b1.csv $L > $O/ncu_b1.log 2>&1; echo "ncu b1 rc $?"
python - <<'PY'
import csv, collections
rows = [ln for ln in open("gpurun_out/r02/ncu_launches_b1.csv") if ln.startswith('"')]
ks = [r for r in csv.DictReader(rows) if r["Metric Name"] == "gpu__time_duration.sum" and "hmv::" in r["Kernel Name"]]
# the first model is 5 views B=1: 70 forwards x 48 kernels; take the last forward of that model
per = 48
first = ks[: 70 * per][-per:]
tot = 0.0
agg = collections.OrderedDict()
for r in first:
    us = float(r["Metric Value"].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r["Metric Unit"], 1.0)
    name = r["Kernel Name"].split("(")[0].replace("void hmv::", "").replace("<unnamed>::", "")[:60]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us; tot += us
print("B=1 (5 views) serialised kernel time %.1f us over %d kernels" % (tot, len(first)))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]): print("  %-62s x%-2d %8.1f us" % (k, v[0], v[1]))
PY
rm -f $O/*.ncu-rep.tmp; ls -la $O | head -40
