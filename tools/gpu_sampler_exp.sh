#!/bin/bash
# does the clock sampler cause the ~70 ms outlier steps?  4 repetitions per variant
mkdir -p gpurun_out
for variant in off full20 clock20 full200; do
  for rep in 1 2 3 4; do
    case $variant in
      off) extra="--no-clocks"; export HMV_BENCH_SAMPLER=full HMV_BENCH_SAMPLER_PERIOD=0.02;;
      full20) extra=""; export HMV_BENCH_SAMPLER=full HMV_BENCH_SAMPLER_PERIOD=0.02;;
      clock20) extra=""; export HMV_BENCH_SAMPLER=clock HMV_BENCH_SAMPLER_PERIOD=0.02;;
      full200) extra=""; export HMV_BENCH_SAMPLER=full HMV_BENCH_SAMPLER_PERIOD=0.2;;
    esac
    timeout 300 python bench.py --steps 40 --warmup 5 --micro-batch 64 --no-cpu-baseline $extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
c=d['clocks'] or {}
print('$variant $rep value %.0f ms %.2f'%(d['value'], d['ms_per_step']), {k: round(v,2) for k,v in d['step_ms'].items()}, 'query_ms_max', c.get('query_ms_max'), 'samples', c.get('samples'))
"
  done
done
