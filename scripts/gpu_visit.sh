#!/bin/bash
# One GPU visit: whole gpu test suite (one pytest process, durations logged), then phase / profile scripts.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout ${TEST_TIMEOUT:-1200} python -m pytest tests -q -m gpu --tb=short --timeout 600 --durations=15 ${PYTEST_ARGS} > gpurun_out/tests_gpu.log 2>&1
  echo "tests exit $?"; tail -30 gpurun_out/tests_gpu.log
fi
for s in ${SCRIPTS}; do
  timeout 600 python scripts/$s.py > gpurun_out/$s.log 2>&1; echo "$s exit $?"; tail -${TAIL:-30} gpurun_out/$s.log
done
