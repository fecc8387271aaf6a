#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -n 4 2>&1 | tail -5 > gpurun_out/lean_pytest.log
cat gpurun_out/lean_pytest.log
SFB_LIB_PATH=$PWD/stable-fast_b200/sfast_b200/libsfb200_trace.so timeout 600 python tests/gemm_latency.py > gpurun_out/gemm_latency_r02b.jsonl 2> gpurun_out/gemm_latency.err; tail -3 gpurun_out/gemm_latency.err
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras 2>gpurun_out/lean_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['ms_per_step'], d['e2e'].get('ms_per_step'))"
timeout 300 python tests/graph_breakdown.py 2 > gpurun_out/breakdown_b2_lean.jsonl 2>gpurun_out/breakdown.err; head -16 gpurun_out/breakdown_b2_lean.jsonl
timeout 600 python tests/gemm_shapes_bench.py lean2 > gpurun_out/gemm_shapes_lean.jsonl 2> gpurun_out/gemm_shapes_lean.err
cut -c1-120 gpurun_out/gemm_shapes_lean.jsonl
