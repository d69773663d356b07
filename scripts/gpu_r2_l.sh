#!/bin/bash
# round-2 visit L: attention backward with 16 compute warps vs 8; full GPU suite
mkdir -p gpurun_out
timeout 300 python scripts/gpu_probe.py attn_bwd perf_bwd > gpurun_out/r2l_probe_wg4.log 2>&1; cat gpurun_out/r2l_probe_wg4.log | tail -16
ADVGRPO_ATTN_BWD_WG=2 timeout 300 python scripts/gpu_probe.py perf_bwd > gpurun_out/r2l_probe_wg2.log 2>&1; tail -4 gpurun_out/r2l_probe_wg2.log
timeout 2400 python -m pytest tests -q -m gpu --timeout 1200 --durations=5 > gpurun_out/r2l_tests.log 2>&1; echo "tests exit $?"; tail -15 gpurun_out/r2l_tests.log
