#!/bin/bash
# round-2 last visit: GPU suite + smoke at head, bench lines for configs 2 / 3 / 5 at N = 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --timeout 900 > gpurun_out/r2l_tests.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/r2l_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2l_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/r2l_smoke.log
for c in 2 3 5; do
timeout 600 python bench.py --config $c --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench_cfg$c.json 2> gpurun_out/r2l_bench_cfg$c.err; echo "bench$c exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r2l_bench_cfg$c.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['phases_ms'], d['clocks']['sm_mhz'])"
done
