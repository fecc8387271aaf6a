#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 1500 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest tests/test_svd_gpu.py -m gpu -x -q -k "tiny_svd_unet_vs_oracle and False" 2>&1 | grep -v CUDAEvent | head -150 > gpurun_out/dbg_sanitizer.log; head -110 gpurun_out/dbg_sanitizer.log
