#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python tests/kernel_checks.py > gpurun_out/kernel_checks.jsonl 2> gpurun_out/kernel_checks.err
echo "checks rc=$?" >> gpurun_out/kernel_checks.jsonl
grep -c '"pass": true' gpurun_out/kernel_checks.jsonl; grep -v '"pass": true' gpurun_out/kernel_checks.jsonl | cut -c1-400
tail -5 gpurun_out/kernel_checks.err
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-roofline"
echo "== default (cluster_k max16, epi8)"; $B 2>>gpurun_out/bench.err | tee gpurun_out/bench_ck16.json | cut -c1-330
echo "== ck max 8"; SFB_CLUSTER_K_MAX=8 $B 2>>gpurun_out/bench.err | tee gpurun_out/bench_ck8.json | cut -c1-330
echo "== ck off"; SFB_CLUSTER_K=0 $B 2>>gpurun_out/bench.err | tee gpurun_out/bench_ck0.json | cut -c1-330
echo "== epi4 ck16"; SFB_LIB_PATH=$PWD/stable-fast_b200/sfast_b200/libsfb200_epi4.so $B 2>>gpurun_out/bench.err | tee gpurun_out/bench_epi4.json | cut -c1-330
echo "== epi4 ck off"; SFB_CLUSTER_K=0 SFB_LIB_PATH=$PWD/stable-fast_b200/sfast_b200/libsfb200_epi4.so $B 2>>gpurun_out/bench.err | tee gpurun_out/bench_epi4_ck0.json | cut -c1-330
echo "== B16 default"; $B --batch 16 --steps 20 2>>gpurun_out/bench.err | tee gpurun_out/bench_b16.json | cut -c1-330
echo "== B16 epi4"; SFB_LIB_PATH=$PWD/stable-fast_b200/sfast_b200/libsfb200_epi4.so $B --batch 16 --steps 20 2>>gpurun_out/bench.err | tee gpurun_out/bench_b16_epi4.json | cut -c1-330
tail -5 gpurun_out/bench.err
timeout 300 python tests/gemm_latency.py > gpurun_out/gemm_latency_epi8.jsonl 2>gpurun_out/gemm_latency.err; cut -c1-420 gpurun_out/gemm_latency_epi8.jsonl
timeout 300 python tests/graph_breakdown.py 2 > gpurun_out/breakdown_b2.jsonl 2>gpurun_out/breakdown.err; head -12 gpurun_out/breakdown_b2.jsonl; tail -2 gpurun_out/breakdown_b2.jsonl
timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -4
