timeout 900 python -m pytest tests -q -m gpu --no-header -x -p no:cacheprovider 2>&1 | tail -5
timeout 600 python bench.py --steps 40 --warmup 5 --no-cpu-baseline > gpurun_out/bench_async.json 2>gpurun_out/bench_async.err; echo rc $?
python -c "
import json
d=json.load(open('gpurun_out/bench_async.json'))
print('value %.0f ms %.2f e2e %.0f e2e_ms %.2f sync_ms %.2f'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['sync_call_ms']), d['roofline']['phase_ms_per_step'])
"
