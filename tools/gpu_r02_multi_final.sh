#!/bin/bash
# end of round 2 (run with gpurun --gpus 8): the default line on 8 GPUs with the final kernels
mkdir -p gpurun_out/r02
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29517"
$T --nproc-per-node 8 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-eager --no-latency > gpurun_out/r02/bench_8gpu.json 2> gpurun_out/r02/bench_8gpu.err; echo "8gpu default rc $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02/bench_8gpu.json")); e = d.get("e2e") or {}
print("N=%d value %.0f poses/s ms/step %.2f | e2e u8 %s fp32 %s | %s" % (d["n_gpus"], d["value"], d["ms_per_step"], e.get("value"), (e.get("fp32_input") or {}).get("value"), d["config"]["workload"]))
PY
