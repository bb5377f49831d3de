#!/bin/bash
# where the tail kernels' MMA thread spends its non-waiting time (issue vs commit)
mkdir -p gpurun_out
HMV_BT_PROF=1 timeout 300 python bench.py --steps 3 --warmup 3 --ramp-seconds 0.5 --no-e2e --no-eager --no-latency --no-cpu-baseline --no-clocks 2> gpurun_out/bt_prof_n.err > /dev/null; grep bt_prof gpurun_out/bt_prof_n.err | head -7
