#!/bin/bash
# compute-sanitizer memcheck of the folded-GroupNorm conv cases (caching allocator off: every tensor its own cudaMalloc)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
timeout 110 compute-sanitizer --tool memcheck --print-limit 6 python tests/kernel_checks.py conv_gn_16_one_block conv_gn_24_ragged_rows conv_gn_40x24_ragged_n conv_gn_concat_pitch conv_gn_bf16 > gpurun_out/memcheck_conv_gn.log 2>&1
grep -c '"pass": true' gpurun_out/memcheck_conv_gn.log; grep -n '"pass"\|Invalid\|ERROR SUMMARY\|at sfb::\|in .*\.cu:' gpurun_out/memcheck_conv_gn.log | head -20
