#!/bin/bash
# early PDL trigger in the non-GEMM kernels + N = 128 tiles for the QKV projection of small passes: the tests that reach them,
# B = 1 latency (against 0.679 / 0.698 ms of the previous build; HMV_NARROW_SMALL=0 for the tile part), a short B = 64 bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu --no-header -rA -p no:cacheprovider -k "fusion or stagewise or chained or known or micro_batching or uint8 or bench_configuration or end_to_end or preprocess" > gpurun_out/pytest_sub.log 2>&1; echo "pytest rc $?"
tail -1 gpurun_out/pytest_sub.log; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_sub.log | head
for e in "HMV_NARROW_SMALL=1" "HMV_NARROW_SMALL=0" "HMV_NO_PDL=1"; do
  echo "== $e"; env $e timeout 200 python tools/bench_latency.py 200 2>&1 | tail -3
done | tee gpurun_out/latency_pdl_ab.txt
timeout 300 python bench.py --steps 30 --warmup 3 --no-e2e --no-eager --no-latency --no-cpu-baseline > gpurun_out/bench_q.json 2>/dev/null
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_q.json")); r = d["roofline"]
print("B=64: value %.0f step median %.3f ms | phases %s" % (d["value"], d["step_ms"]["median"], r.get("phase_ms_per_step")))
PY
