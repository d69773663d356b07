#!/bin/bash
# round-2 visit S: score-head / discriminator-step kernels, native VAE decoder, D-step parity
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_heads_gpu.py -q -m gpu --timeout 600 -x > gpurun_out/r2s_heads.log 2>&1; echo "heads exit $?"; tail -15 gpurun_out/r2s_heads.log
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_rewards_gpu.py tests/test_tensorcore_gpu.py -q -m gpu --timeout 600 -k "layer_norm or vae or conv or dino or pickscore or gemm" > gpurun_out/r2s_kernels.log 2>&1; echo "kernels exit $?"; tail -15 gpurun_out/r2s_kernels.log
timeout 900 python -m pytest tests/test_fullsize_gpu.py tests/test_pipeline_gpu.py -q -m gpu --timeout 600 -k "dino or pickscore or discriminator or smoke" -s > gpurun_out/r2s_dstep.log 2>&1; echo "dstep exit $?"; tail -15 gpurun_out/r2s_dstep.log
