#!/bin/bash
# end of round 2 (gpurun --gpus 8): BASELINE config 3 (B = 4096, batch-sharded) and config 4 (8 views) on the final kernels
mkdir -p gpurun_out/r02
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29517"
$T --nproc-per-node 8 bench.py --gpus 8 --steps 5 --warmup 3 --batch 512 --no-e2e --no-cpu-baseline --no-eager --no-latency > gpurun_out/r02/bench_8gpu_b4096.json 2> gpurun_out/r02/bench_8gpu_b4096.err; echo "8gpu B=4096 rc $?"
$T --nproc-per-node 8 bench.py --gpus 8 --steps 20 --warmup 5 --views 8 --no-cpu-baseline --no-eager --no-latency > gpurun_out/r02/bench_8gpu_views8.json 2> gpurun_out/r02/bench_8gpu_views8.err; echo "8gpu 8 views rc $?"
python - <<'PY'
import json
for f in ("bench_8gpu_b4096", "bench_8gpu_views8"):
    d = json.load(open("gpurun_out/r02/%s.json" % f)); e = d.get("e2e") or {}
    print(f, "N=%d value %.0f poses/s ms/step %.2f | e2e u8 %s | %s" % (d["n_gpus"], d["value"], d["ms_per_step"], e.get("value"), d["config"]["workload"]))
PY
