#!/bin/bash
# round-2 visit I: pair kernel (two query tiles per CTA in antiphase) correctness + perf
mkdir -p gpurun_out
timeout 300 python scripts/gpu_probe.py attn_pair perf_attn > gpurun_out/r2i_probe.log 2>&1; echo "probe exit $?"; cat gpurun_out/r2i_probe.log | tail -90
