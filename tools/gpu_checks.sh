#!/bin/bash
# Bring-up run on the GPU box: each phase in its own process (a faulting kernel poisons its context only).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for phase in conv_fp32 conv_bf16 model_fp32 model_bf16; do
  timeout 600 python tools/diag.py $phase > gpurun_out/diag_${phase}.log 2>&1
  echo "phase $phase exit $?" >> gpurun_out/phases.txt
done
timeout 1500 python -m pytest tests -q -m gpu -x --no-header -rA -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/phases.txt
tail -5 gpurun_out/pytest_gpu.log
cat gpurun_out/phases.txt
