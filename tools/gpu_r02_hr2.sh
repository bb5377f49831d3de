#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider -k "hrnet" > gpurun_out/pytest_gpu_hr.log 2>&1; echo "pytest rc $?"; tail -2 gpurun_out/pytest_gpu_hr.log
python bench.py --backbone hrnet --steps 20 --no-cpu-baseline --no-eager --no-e2e 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('hrnet value %.1f ms/step %.2f' % (d['value'], d['ms_per_step']), d['roofline']['phase_ms_per_step'])"
