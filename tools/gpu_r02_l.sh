#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -k "stagewise or known or chained or end_to_end or micro_batching" > gpurun_out/pytest_gpu_l.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/pytest_gpu_l.log
grep -E "^(FAILED|ERROR|E  )" gpurun_out/pytest_gpu_l.log | cut -c1-300
python tools/bench_latency.py 300
python bench.py --steps 30 --no-e2e --no-eager --no-latency --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('value', d['value'], d['step_ms'], d['roofline']['phase_ms_per_step'])"
