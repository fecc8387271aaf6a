#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python tests/kernel_checks.py > gpurun_out/kernel_checks.jsonl 2> gpurun_out/kernel_checks.err
echo "checks rc=$?" >> gpurun_out/kernel_checks.jsonl
grep -c '"pass": true' gpurun_out/kernel_checks.jsonl; grep -v '"pass": true' gpurun_out/kernel_checks.jsonl | cut -c1-300
timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -8
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --dump-ops gpurun_out/ops_b2.jsonl > gpurun_out/bench.json 2>gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json
timeout 300 python bench.py --steps 20 --warmup 3 --batch 16 --no-cpu-baseline --dump-ops gpurun_out/ops_b16.jsonl > gpurun_out/bench_b16.json 2>>gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_b16.json
SFB_GN_EPILOGUE=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-roofline 2>>gpurun_out/bench.err | cut -c1-200
timeout 300 python tests/graph_breakdown.py 2 > gpurun_out/breakdown_b2.jsonl 2>gpurun_out/breakdown.err; head -14 gpurun_out/breakdown_b2.jsonl; tail -2 gpurun_out/breakdown_b2.jsonl
tail -3 gpurun_out/bench.err
