#!/bin/bash
# round-2 visit M: attention backward with two issuer warps (NWG 4 and 2), HBM-kernel table + ncu, GEMM ncu (traffic)
mkdir -p gpurun_out
timeout 300 python scripts/gpu_probe.py attn_bwd perf_bwd > gpurun_out/r2m_probe_wg4.log 2>&1; cat gpurun_out/r2m_probe_wg4.log | tail -16
ADVGRPO_ATTN_BWD_WG=2 timeout 300 python scripts/gpu_probe.py attn_bwd perf_bwd > gpurun_out/r2m_probe_wg2.log 2>&1; tail -16 gpurun_out/r2m_probe_wg2.log
timeout 600 python -m pytest tests/test_tensorcore_gpu.py tests/test_kernels_gpu.py -q -m gpu -x --timeout 600 -k "bwd or sde or autograd or replay" > gpurun_out/r2m_tests.log 2>&1; echo "tests exit $?"; tail -5 gpurun_out/r2m_tests.log
timeout 600 python scripts/profile_hbm_kernels.py > gpurun_out/r2m_hbm_kernels.log 2>&1; echo "hbm exit $?"; cat gpurun_out/r2m_hbm_kernels.log
timeout 900 ncu --set full --clock-control none -k regex:'sde_|group_advantage|grpo_clip|ln_modulate|qk_norm' -o gpurun_out/r2m_hbm python scripts/profile_hbm_kernels.py --once > gpurun_out/r2m_ncu_hbm.log 2>&1; echo "ncu hbm exit $?"; tail -3 gpurun_out/r2m_ncu_hbm.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 2 -c 1 -o gpurun_out/r2m_gemm python scripts/profile_gemm.py > gpurun_out/r2m_ncu_gemm.log 2>&1; echo "ncu gemm exit $?"; tail -2 gpurun_out/r2m_ncu_gemm.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 1 -c 1 -o gpurun_out/r2m_attn_bwd python scripts/profile_attn.py > gpurun_out/r2m_ncu_bwd.log 2>&1; echo "ncu bwd exit $?"; tail -2 gpurun_out/r2m_ncu_bwd.log
