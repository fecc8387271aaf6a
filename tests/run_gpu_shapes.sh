#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/gemm_shapes.jsonl
for lib in libsfb200.so libsfb200_v1.so libsfb200_v2.so libsfb200_v3.so; do
  SFB_LIB_PATH=$PWD/stable-fast_b200/sfast_b200/$lib timeout 300 python tests/gemm_shapes_bench.py $lib >> gpurun_out/gemm_shapes.jsonl 2>> gpurun_out/gemm_shapes.err
done
cat gpurun_out/gemm_shapes.jsonl
