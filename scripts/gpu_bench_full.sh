#!/bin/bash
# default bench (with cpu_baseline), reference arm, smoke, and the ncu launch list of a short bench run
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref exit $?"
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"
tail -3 gpurun_out/bench_default.err gpurun_out/bench_reference.err gpurun_out/smoke.log
cat gpurun_out/bench_default.json gpurun_out/bench_reference.json
