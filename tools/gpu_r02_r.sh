#!/bin/bash
# is the enqueueing thread throttled by the container's CPU quota?  cgroup counters before / after bench runs
mkdir -p gpurun_out
echo "nproc $(nproc)"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /sys/fs/cgroup/cpu/cpu.cfs_quota_us /sys/fs/cgroup/cpu/cpu.cfs_period_us 2>/dev/null
stat() { grep -E "nr_periods|nr_throttled|throttled" /sys/fs/cgroup/cpu.stat /sys/fs/cgroup/cpu/cpu.stat 2>/dev/null | tr '\n' ' '; echo; }
Q="--steps 20 --warmup 3 --no-e2e --no-eager --no-latency --no-cpu-baseline --no-clocks"
for i in 1 2 3 4 5 6 7 8; do
  stat
  HMV_PAIR_MINK=512 timeout 300 python bench.py $Q > gpurun_out/bench_v.json 2>/dev/null
  python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_v.json"))
print("value %.0f step median %.3f max %.3f | enqueue %.2f ms/step | slowest %s" % (d["value"], d["step_ms"]["median"], d["step_ms"]["max"], d["step_ms"]["cpu_enqueue"], d["step_ms"]["cpu_enqueue_slowest"]))
PY
done
stat
