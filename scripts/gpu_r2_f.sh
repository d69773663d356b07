#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --timeout 1200 --durations=8 > gpurun_out/r2f_tests.log 2>&1; echo "tests exit $?"; tail -45 gpurun_out/r2f_tests.log
