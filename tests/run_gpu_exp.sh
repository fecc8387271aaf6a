#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python tests/kernel_checks.py > gpurun_out/kernel_checks.jsonl 2> gpurun_out/kernel_checks.err
echo "checks rc=$?" >> gpurun_out/kernel_checks.jsonl
grep -c '"pass": true' gpurun_out/kernel_checks.jsonl; grep -v '"pass": true' gpurun_out/kernel_checks.jsonl | cut -c1-400
tail -5 gpurun_out/kernel_checks.err
timeout 1500 python -m pytest tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -4
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-roofline"
echo "== default (upconv on)"; $B 2>>gpurun_out/bench.err | tee gpurun_out/bench_up1.json | cut -c1-330
echo "== upconv off"; SFB_UPCONV=0 $B 2>>gpurun_out/bench.err | tee gpurun_out/bench_up0.json | cut -c1-330
echo "== B16 default"; $B --batch 16 --steps 20 2>>gpurun_out/bench.err | tee gpurun_out/bench_b16.json | cut -c1-330
echo "== B16 upconv off"; SFB_UPCONV=0 $B --batch 16 --steps 20 2>>gpurun_out/bench.err | tee gpurun_out/bench_b16_up0.json | cut -c1-330
tail -5 gpurun_out/bench.err
