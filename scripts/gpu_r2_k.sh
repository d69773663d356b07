#!/bin/bash
# round-2 visit K: full GPU suite with the pair kernel as the D=64 default + bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x --timeout 1200 --durations=5 > gpurun_out/r2k_tests.log 2>&1; echo "tests exit $?"; tail -15 gpurun_out/r2k_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench exit $?"; tail -c 3500 gpurun_out/r2k_bench.json; tail -3 gpurun_out/r2k_bench.err
