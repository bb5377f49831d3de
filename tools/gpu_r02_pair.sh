#!/bin/bash
# cta_group::2 pair MMAs in the conv GEMM kernel: correctness on the clustered geometry classes, then A/B
mkdir -p gpurun_out
HMV_PAIR=1 timeout 600 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -x -k "conv_bn_act and bf16" > gpurun_out/pair_test.log 2>&1; echo "pair conv tests rc $?"; tail -3 gpurun_out/pair_test.log; grep -E "^E  " gpurun_out/pair_test.log | head -5 | cut -c1-300
HMV_PAIR=1 timeout 600 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -x -k "steps_teacher_forced and bf16 and not hrnet or stagewise_teacher_forced and bf16-5-True" > gpurun_out/pair_test2.log 2>&1; echo "pair step tests rc $?"; tail -2 gpurun_out/pair_test2.log; grep -E "^E  " gpurun_out/pair_test2.log | head -5 | cut -c1-300
Q="--steps 30 --warmup 3 --no-e2e --no-eager --no-latency --no-cpu-baseline"
for e in HMV_PAIR=0 HMV_PAIR=1 HMV_PAIR=0 HMV_PAIR=1; do
  env $e timeout 300 python bench.py $Q > gpurun_out/bench_v.json 2>/dev/null
  python - "$e" <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_v.json")); r = d["roofline"]
cl = {c["kernel"]: c for c in r["classes"]}
print("%-12s step median %.3f | " % (sys.argv[1], d["step_ms"]["median"]) + " ".join("%s %.4f" % (k[7:] if k.startswith("layer3") else k, cl[k]["ms_per_launch"]) for k in ("layer3.x.conv2", "layer3.x.downsample", "layer3.x.conv1", "pose_net.0", "sample_nets.0", "fusion.x.qkv") if k in cl))
PY
done
