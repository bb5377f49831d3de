#!/bin/bash
# GPU tests + bench (optionally A/B against an env switch: AB_ENV="HMV_FUSION_UNFUSED=1")
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc $?"
tail -3 gpurun_out/pytest_gpu.log
grep -E "^(FAILED|E  )|teacher-forced|flip rate" gpurun_out/pytest_gpu.log | cut -c1-250
run_bench() {
  timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$1.json 2>gpurun_out/bench_$1.err; echo "bench $1 rc $?"
  cp gpurun_out/tc_launches.csv gpurun_out/tc_launches_$1.csv
  python -c "
import json
d=json.load(open('gpurun_out/bench_$1.json'))
print('$1: value %.0f ms %.2f (median %.2f) e2e %.0f e2e_ms %.2f sync_ms %.2f launches/step %.0f'%(d['value'], d['ms_per_step'], d['step_ms']['median'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['sync_call_ms'], d['gpu_launches']/d['steps']), {k: round(v,3) for k,v in d['roofline']['phase_ms_per_step'].items()})
"
}
run_bench main
if [ -n "$AB_ENV" ]; then env $AB_ENV bash -c "$(declare -f run_bench); run_bench ab"; fi
