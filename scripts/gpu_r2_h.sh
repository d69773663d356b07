#!/bin/bash
# round-2 visit H (re-entry): ncu --set full of the default attention forward (quad) and backward, probe perf
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python scripts/gpu_probe.py perf_attn perf_bwd > gpurun_out/r2h_probe.log 2>&1; echo "probe exit $?"; grep -i "attn\|sdpa" gpurun_out/r2h_probe.log
VARIANT=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_quad -s 2 -c 1 -o gpurun_out/r2h_attn_quad python scripts/profile_attn_fwd.py > gpurun_out/r2h_ncu.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/r2h_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 1 -c 1 -o gpurun_out/r2h_attn_bwd python scripts/profile_attn.py > gpurun_out/r2h_ncu_bwd.log 2>&1; echo "ncu bwd exit $?"; tail -3 gpurun_out/r2h_ncu_bwd.log
