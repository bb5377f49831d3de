#!/bin/bash
# ncu --set full captures of one representative launch of every kernel class (run on the GPU box).
# The same command lines are first run without ncu (exit code checked) as the profiling recipe requires.
mkdir -p gpurun_out/ncu
B="python bench.py --steps 1 --warmup 3 --ramp-seconds 0 --micro-batch 64 --no-cpu-baseline --no-clocks"
$B > gpurun_out/ncu/plain_bench.log 2>&1 || { echo "plain bench failed"; exit 1; }
cap() {  # name, regex, skip, cmd...
  name=$1; regex=$2; skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$regex" -s $skip -c 1 -f -o gpurun_out/ncu/$name "$@" > gpurun_out/ncu/$name.log 2>&1
  echo "ncu $name exit $?"
}
for c in l3.conv2 l3.conv3 l3.conv1 l1.conv2 l2.conv2; do
  python tools/bench_conv.py 160 $c 1 > gpurun_out/ncu/plain_$c.log 2>&1 && cap conv_$c conv_gemm_tc 1 python tools/bench_conv.py 160 $c 1
done
cap stem_pool stem_pool_kernel 3 $B
cap attention_mma attention_mma_kernel 15 $B
cap layernorm layernorm_kernel 30 $B
cap gcn_l1 gcn_l1_kernel 3 $B
cap softargmax softargmax_kernel 3 $B
cap tokens tokens_kernel 3 $B
cap sample_gather sample_gather_kernel 3 $B
ls -la gpurun_out/ncu | head -40
