#!/bin/bash
# round-2 visit B: quad-layout attention forward (v2 pipeline) correctness + perf + ncu
mkdir -p gpurun_out
timeout 600 python scripts/gpu_probe.py attn_quad perf_attn > gpurun_out/r2b_probe_attn.log 2>&1; echo "probe exit $?"; grep -v "variant=1[68]" gpurun_out/r2b_probe_attn.log | tail -40
VARIANT=17 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_quad -s 2 -c 1 -o gpurun_out/r2b_attn_quad17 python scripts/profile_attn_fwd.py > gpurun_out/r2b_ncu.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/r2b_ncu.log
