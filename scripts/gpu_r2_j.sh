#!/bin/bash
# round-2 visit J: pair kernel perf + ncu --set full
mkdir -p gpurun_out
timeout 300 python scripts/gpu_probe.py perf_attn > gpurun_out/r2j_probe.log 2>&1; echo "probe exit $?"; grep "S=1229\|S=4301" gpurun_out/r2j_probe.log
VARIANT=21 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_pair -s 2 -c 1 -o gpurun_out/r2j_attn_pair python scripts/profile_attn_fwd.py > gpurun_out/r2j_ncu.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/r2j_ncu.log
