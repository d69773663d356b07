#!/bin/bash
# round-2 visit N: fast grpo advantage kernel + warp-parallel sde schedule lookup + fast-math Box-Muller: tests, HBM table, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_cabi.py tests/test_fullsize_gpu.py -q -m gpu --timeout 600 > gpurun_out/r2n_tests.log 2>&1; echo "tests exit $?"; tail -8 gpurun_out/r2n_tests.log
timeout 600 python scripts/profile_hbm_kernels.py > gpurun_out/r2n_hbm_kernels.log 2>&1; echo "hbm exit $?"; cat gpurun_out/r2n_hbm_kernels.log
timeout 900 ncu --set full --clock-control none -k regex:'sde_|group_advantage|grpo_clip|ln_modulate|qk_norm' -o gpurun_out/r2n_hbm python scripts/profile_hbm_kernels.py --once > gpurun_out/r2n_ncu_hbm.log 2>&1; echo "ncu hbm exit $?"; tail -3 gpurun_out/r2n_ncu_hbm.log
