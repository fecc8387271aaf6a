#!/bin/bash
# Two-epilogue-group persistent GEMM: correctness (persist_* kernel checks) then per-shape A/B.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "persist or gemm or conv or geglu or qkv" 2>&1 | tail -5 > gpurun_out/g2_pytest.log
cat gpurun_out/g2_pytest.log
timeout 600 python tests/gemm_shapes_bench.py g2 > gpurun_out/gemm_shapes_g2.jsonl 2> gpurun_out/gemm_shapes_g2.err
cat gpurun_out/gemm_shapes_g2.jsonl; tail -3 gpurun_out/gemm_shapes_g2.err
