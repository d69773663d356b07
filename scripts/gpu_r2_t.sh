#!/bin/bash
# round-2 visit T: full GPU suite + smoke + bench lines (config 2 / 3 / 5 at N=1) after the native discriminator step / VAE
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 -x > gpurun_out/r2t_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/r2t_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2t_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r2t_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_bench_cfg2.json 2> gpurun_out/r2t_bench_cfg2.err; echo "bench2 exit $?"; cut -c1-400 gpurun_out/r2t_bench_cfg2.json
timeout 600 python bench.py --config 3 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_bench_cfg3.json 2> gpurun_out/r2t_bench_cfg3.err; echo "bench3 exit $?"; cut -c1-300 gpurun_out/r2t_bench_cfg3.json
timeout 600 python bench.py --config 5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2t_bench_cfg5.json 2> gpurun_out/r2t_bench_cfg5.err; echo "bench5 exit $?"; cut -c1-300 gpurun_out/r2t_bench_cfg5.json
