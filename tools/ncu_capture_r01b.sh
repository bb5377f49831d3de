#!/bin/bash
# ncu captures of the kernels added in the second half of round 1 (fused bottleneck tail, fused fusion layer) plus a
# launch list of one whole step.  Every command is first run without ncu (exit code checked).
mkdir -p gpurun_out/ncu
B="python bench.py --steps 1 --warmup 3 --ramp-seconds 0 --no-cpu-baseline --no-clocks --no-e2e"
$B > gpurun_out/ncu/plain_bench.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/ncu/plain_bench.log; exit 1; }
cap() {  # name, regex, skip, cmd...
  name=$1; regex=$2; skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s $skip -c 1 -f -o gpurun_out/ncu/$name "$@" > gpurun_out/ncu/$name.log 2>&1
  echo "ncu $name exit $?"
}
cap tail_p64 bottleneck_tail_kernel 7 $B
cap tail_p128 bottleneck_tail_kernel 11 $B
cap fusion_block_l0 fusion_block_kernel 5 $B
cap fusion_block_l3 fusion_block_kernel 8 $B
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 140 --csv --log-file gpurun_out/launches_r01b.csv $B > gpurun_out/ncu/launchlist.log 2>&1
echo "launch list exit $?"
ls -la gpurun_out/ncu | head -40
