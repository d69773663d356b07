#!/bin/bash
mkdir -p gpurun_out
for t in 0 1; do
echo "== ADVGRPO_ATTN_TMA_OUT=$t"
ADVGRPO_ATTN_TMA_OUT=$t timeout 600 python scripts/gpu_probe.py attn_split perf_attn > gpurun_out/r2c_probe_attn_$t.log 2>&1; echo "probe exit $?"; grep -v "S=1024 H\|S=1370\|variant 1[68]\|variant 3" gpurun_out/r2c_probe_attn_$t.log | tail -12
done
timeout 300 python scripts/trace_attn.py > gpurun_out/r2c_trace.log 2>&1; echo "trace exit $?"; sed -n 3,20p gpurun_out/r2c_trace.log
