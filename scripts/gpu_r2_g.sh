#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --timeout 1200 > gpurun_out/r2g_tests.log 2>&1; echo "tests exit $?"; tail -12 gpurun_out/r2g_tests.log
for c in 3 5; do
timeout 900 python bench.py --config $c --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_cfg$c.json 2> gpurun_out/r2g_bench_cfg$c.err; echo "bench cfg $c exit $?"; tail -c 1500 gpurun_out/r2g_bench_cfg$c.json; tail -5 gpurun_out/r2g_bench_cfg$c.err
done
timeout 1500 python bench.py --config 4 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench_cfg4.json 2> gpurun_out/r2g_bench_cfg4.err; echo "bench cfg 4 exit $?"; tail -c 1500 gpurun_out/r2g_bench_cfg4.json; tail -5 gpurun_out/r2g_bench_cfg4.err
