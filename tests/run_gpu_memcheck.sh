#!/bin/bash
# Memory-safety sweep: every kernel check and the eager tiny-model tests under compute-sanitizer with
# PyTorch's caching allocator off (each tensor its own cudaMalloc, so any out-of-bounds access is seen).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
timeout 1500 compute-sanitizer --tool memcheck --print-limit 6 python tests/kernel_checks.py > gpurun_out/memcheck_kernels.log 2>&1
grep -c '"pass": true' gpurun_out/memcheck_kernels.log; grep -c '"pass": false' gpurun_out/memcheck_kernels.log
grep -n "Invalid\|ERROR SUMMARY\|at sfb::\|in .*\.cu:" gpurun_out/memcheck_kernels.log | head -30
timeout 1500 compute-sanitizer --tool memcheck --print-limit 6 python -m pytest tests/test_svd_gpu.py tests/test_unet_gpu.py tests/test_vae_gpu.py -m gpu -q -k "${MEMCHECK_MODELS:-(tiny and False) or rectangular or odd}" > gpurun_out/memcheck_models.log 2>&1
grep -n "Invalid\|ERROR SUMMARY\|at sfb::\|in .*\.cu:\|passed\|failed" gpurun_out/memcheck_models.log | head -30
