#!/bin/bash
# A/B of the 2-CTA cluster + multicast-weights variant of the conv kernel: parity of the affected geometries, per-layer
# micro-benchmark, whole-model bench
mkdir -p gpurun_out
timeout 200 python -m pytest tests -q -m gpu --no-header -x -p no:cacheprovider -k "conv_bn_act and bf16" 2>&1 | tail -3
for c in 1 0; do
  echo "== HMV_CLUSTER=$c"
  for l in l3.0.conv1 l3.0.down l3.conv1 l3.conv2 l3.conv3 pose0; do HMV_CLUSTER=$c timeout 120 python tools/bench_conv.py 320 $l 10 | tail -n +2 | grep -E "^$l " ; done
done
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_cluster1.json 2>/dev/null; cp gpurun_out/tc_launches.csv gpurun_out/tc_launches_cluster1.csv
HMV_CLUSTER=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_cluster0.json 2>/dev/null
python - <<'PY'
import json
for f in ("bench_cluster1", "bench_cluster0"):
    d = json.load(open(f"gpurun_out/{f}.json"))
    print(f, "value %.0f ms %.2f median %.2f" % (d["value"], d["ms_per_step"], d["step_ms"]["median"]), {k: round(v, 3) for k, v in d["roofline"]["phase_ms_per_step"].items()})
PY
