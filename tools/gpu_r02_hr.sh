#!/bin/bash
# round 2: HRNet bring-up
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --no-header -rA -p no:cacheprovider -k "hrnet or known_answers or golden_fixtures" > gpurun_out/pytest_gpu_hr.log 2>&1; echo "pytest rc $?"
tail -3 gpurun_out/pytest_gpu_hr.log
grep -E "^(FAILED|ERROR|E  )|\[hrnet" gpurun_out/pytest_gpu_hr.log | cut -c1-700 | head -40
