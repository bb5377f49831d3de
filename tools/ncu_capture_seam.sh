#!/bin/bash
# ncu --set full of one launch of the layer3 seam kernel (the bench command is first run without ncu)
mkdir -p gpurun_out/ncu
B="python bench.py --steps 1 --warmup 3 --ramp-seconds 0 --no-cpu-baseline --no-clocks --no-e2e"
$B > gpurun_out/ncu/plain_bench.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/ncu/plain_bench.log; exit 1; }
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:bottleneck_next_kernel" -s 6 -c 1 -f -o gpurun_out/ncu/seam_l3 $B > gpurun_out/ncu/seam_l3.log 2>&1
echo "ncu seam exit $?"
