#!/bin/bash
# round-2 final validation at head: GPU suite, smoke, driver-form bench (both arms)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 --durations=8 > gpurun_out/r2f_tests.log 2>&1; echo "tests exit $?"; tail -14 gpurun_out/r2f_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2f_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r2f_smoke.log
timeout 900 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench exit $?"; cut -c1-300 gpurun_out/r2f_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 --cpu-budget 60 > gpurun_out/r2f_bench_ref.json 2> gpurun_out/r2f_bench_ref.err; echo "ref exit $?"; cut -c1-300 gpurun_out/r2f_bench_ref.json
