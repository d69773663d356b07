#!/bin/bash
# round-2 visit A: quad-layout attention forward correctness + perf + ncu, GEMM tail-split probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python scripts/gpu_probe.py attn_quad perf_attn > gpurun_out/r2a_probe_attn.log 2>&1; echo "probe exit $?"; tail -60 gpurun_out/r2a_probe_attn.log
timeout 300 python scripts/probe_gemm_tail.py > gpurun_out/r2a_gemm_tail.log 2>&1; echo "tail exit $?"; cat gpurun_out/r2a_gemm_tail.log
VARIANT=17 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_quad -s 2 -c 1 -o gpurun_out/r2a_attn_quad17 python scripts/profile_attn_fwd.py > gpurun_out/r2a_ncu.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/r2a_ncu.log
