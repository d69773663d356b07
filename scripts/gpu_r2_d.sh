#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tensorcore_gpu.py -q -m gpu -x --timeout 600 2>&1 | tail -5
timeout 600 python scripts/gpu_probe.py perf > gpurun_out/r2d_probe_perf.log 2>&1; echo "probe exit $?"; grep "gemm\|attn" gpurun_out/r2d_probe_perf.log | grep -v "1cta\|bn256" | tail -24
