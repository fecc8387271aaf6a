#!/bin/bash
# Lean STORE epilogue + weight prefetch before the PDL wait: correctness, per-shape A/B, B=2 step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -n 4 2>&1 | tail -5 > gpurun_out/lean_pytest.log
cat gpurun_out/lean_pytest.log
timeout 600 python tests/gemm_shapes_bench.py lean > gpurun_out/gemm_shapes_lean.jsonl 2> gpurun_out/gemm_shapes_lean.err
cut -c1-200 gpurun_out/gemm_shapes_lean.jsonl
for kb in 0 12 999; do
  SFB_GEMM_DEEP_MIN_KB=$kb timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras 2>gpurun_out/lean_bench_$kb.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('deep_min_kb=$kb', d['ms_per_step'], d['e2e']['ms_per_step'] if 'ms_per_step' in d['e2e'] else d['e2e'])"
done
SFB_PDL=0 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nopdl', d['ms_per_step'])"
timeout 300 python tests/graph_breakdown.py 2 > gpurun_out/breakdown_b2_lean.jsonl 2>gpurun_out/breakdown.err; head -12 gpurun_out/breakdown_b2_lean.jsonl
