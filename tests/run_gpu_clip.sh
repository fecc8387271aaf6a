#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -n 4 -k "patchify or clip or embed or causal or act" 2>&1 | tail -8 > gpurun_out/clip_kernels.log; cat gpurun_out/clip_kernels.log
timeout 900 python -m pytest tests/test_clip_gpu.py -m gpu -q -s -k "vision or vit" 2>&1 | grep -v Warning | tail -40 > gpurun_out/clip_pytest.log; cat gpurun_out/clip_pytest.log
