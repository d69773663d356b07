#!/bin/bash
# One GPU visit: probes, then the gpu test suite per file (separate processes), logs into gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python scripts/gpu_probe.py "$@" > gpurun_out/probe.log 2>&1
for f in ${TESTS:-tests/test_kernels_gpu.py tests/test_tensorcore_gpu.py tests/test_mmdit_gpu.py tests/test_pipeline_gpu.py tests/test_rewards_gpu.py}; do
  timeout 900 python -m pytest $f -q -m gpu --tb=short --timeout 300 2>&1 | tail -120 > gpurun_out/$(basename $f .py).log
done
tail -50 gpurun_out/probe.log
tail -25 gpurun_out/test_*.log
