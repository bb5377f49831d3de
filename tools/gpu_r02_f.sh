#!/bin/bash
mkdir -p gpurun_out
P="--steps 3 --warmup 3 --ramp-seconds 0.5 --no-e2e --no-eager --no-latency --no-cpu-baseline --no-clocks"
HMV_BN_PROF=1 timeout 300 python bench.py $P 2> gpurun_out/bn_prof_f.err > /dev/null; grep bn_prof gpurun_out/bn_prof_f.err | head -2
