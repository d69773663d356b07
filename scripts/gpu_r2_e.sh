#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --timeout 900 --durations=8 > gpurun_out/r2e_tests.log 2>&1; echo "tests exit $?"; tail -25 gpurun_out/r2e_tests.log
timeout 300 python scripts/gpu_probe.py perf_bwd > gpurun_out/r2e_probe_bwd.log 2>&1; tail -4 gpurun_out/r2e_probe_bwd.log
timeout 300 python scripts/bench_conv.py > gpurun_out/r2e_bench_conv.log 2>&1; tail -12 gpurun_out/r2e_bench_conv.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench exit $?"; tail -c 3000 gpurun_out/r2e_bench.json
