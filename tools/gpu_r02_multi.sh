#!/bin/bash
# round 2, multi-GPU lines (run with gpurun --gpus 8): BASELINE configs 3 (large batch, batch-sharded) and 4 (8 views), default e2e scaling
mkdir -p gpurun_out/r02
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29517"
$T --nproc-per-node 8 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02/bench_8gpu.json 2> gpurun_out/r02/bench_8gpu.err; echo "8gpu default rc $?"
$T --nproc-per-node 8 bench.py --gpus 8 --steps 5 --warmup 3 --batch 512 --no-e2e > gpurun_out/r02/bench_8gpu_b4096.json 2> gpurun_out/r02/bench_8gpu_b4096.err; echo "8gpu B=4096 rc $?"
$T --nproc-per-node 8 bench.py --gpus 8 --steps 20 --warmup 5 --views 8 > gpurun_out/r02/bench_8gpu_views8.json 2> gpurun_out/r02/bench_8gpu_views8.err; echo "8gpu 8 views rc $?"
$T --nproc-per-node 4 bench.py --gpus 4 --steps 10 --warmup 3 --batch 256 --no-e2e > gpurun_out/r02/bench_4gpu_b1024.json 2> gpurun_out/r02/bench_4gpu_b1024.err; echo "4gpu B=1024 rc $?"
$T --nproc-per-node 2 bench.py --gpus 2 --steps 10 --warmup 3 --batch 128 --no-e2e > gpurun_out/r02/bench_2gpu_b256.json 2> gpurun_out/r02/bench_2gpu_b256.err; echo "2gpu B=256 rc $?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02/bench_*gpu*.json")):
    try: d = json.load(open(f))
    except Exception as e: print(f, "unreadable", e); continue
    e = d.get("e2e") or {}
    print(f, "N=%d value %.0f poses/s ms/step %.2f | e2e u8 %s fp32 %s | %s" % (d["n_gpus"], d["value"], d["ms_per_step"], e.get("value"), (e.get("fp32_input") or {}).get("value"), d["config"]["workload"]))
PY
