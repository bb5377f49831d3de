#!/bin/bash
# round-1 second-half measurement set: B=1 latency A/B, launch list with DRAM bytes, default bench, reference arm
mkdir -p gpurun_out
for mode in fused unfused_block unfused_all; do
  case $mode in
    fused) E="";;
    unfused_block) E="HMV_FUSION_UNFUSED=1";;
    unfused_all) E="HMV_FUSION_UNFUSED=1 HMV_FUSE_TAIL=0";;
  esac
  env $E timeout 300 python tools/bench_latency.py 300 > gpurun_out/latency_$mode.jsonl 2>gpurun_out/latency_$mode.err
  echo "latency $mode rc $?"; cat gpurun_out/latency_$mode.jsonl
done
B="python bench.py --steps 1 --warmup 3 --ramp-seconds 0 --no-cpu-baseline --no-clocks --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 150 -c 140 --csv \
   --log-file gpurun_out/launches_dram_r01b.csv $B > gpurun_out/launches_dram.log 2>&1
echo "launch list rc $?"
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc $?"
cat gpurun_out/bench_default.json | cut -c1-1500
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.json 2>/dev/null; echo "ref rc $?"
cat gpurun_out/bench_ref.json | cut -c1-400
