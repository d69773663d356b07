#!/bin/bash
# round-2 visit U: short-sequence attention (fwd + bwd) for the trainable CLIP blocks, D-step parity, config-5 bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_heads_gpu.py -q -m gpu --timeout 600 -x -k "attention_small" > gpurun_out/r2u_small.log 2>&1; echo "small exit $?"; tail -8 gpurun_out/r2u_small.log
timeout 900 python -m pytest tests/test_fullsize_gpu.py tests/test_pipeline_gpu.py tests/test_rewards_gpu.py -q -m gpu --timeout 600 -k "pickscore or discriminator or smoke or cotrain" > gpurun_out/r2u_dstep.log 2>&1; echo "dstep exit $?"; tail -8 gpurun_out/r2u_dstep.log
timeout 600 python bench.py --config 5 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/r2u_bench_cfg5.json 2> gpurun_out/r2u_bench_cfg5.err; echo "bench5 exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/r2u_bench_cfg5.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['phases_ms'], d['clocks'])"
timeout 300 python scripts/profile_dstep.py > gpurun_out/r2u_dstep_profile.log 2>&1; echo "profile exit $?"; tail -40 gpurun_out/r2u_dstep_profile.log
