#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -s -k "128_latent or sdxl_unet_full" 2>&1 | grep -v "^$" | tail -12
B="timeout 600 python bench.py --no-cpu-baseline"
echo "== SDXL B=8 128 bf16"; $B --model sdxl --batch 8 --size 128 --dtype bf16 --steps 10 --warmup 3 2>>gpurun_out/bench.err | tee gpurun_out/bench_sdxl_b8.json | cut -c1-2600
echo "== SD15 B=8 128 fp16"; $B --batch 8 --size 128 --steps 10 --warmup 3 2>>gpurun_out/bench.err | tee gpurun_out/bench_sd15_128_b8.json | cut -c1-2600
echo "== SD15 B=8 64 fp16"; $B --batch 8 --steps 20 --warmup 3 --no-roofline 2>>gpurun_out/bench.err | tee gpurun_out/bench_sd15_b8.json | cut -c1-400
echo "== SD15 B=1"; $B --batch 1 --steps 30 --warmup 5 --no-roofline 2>>gpurun_out/bench.err | tee gpurun_out/bench_sd15_b1.json | cut -c1-400
echo "== stages3 B2"; SFB_GEMM_STAGES=3 $B --steps 30 --warmup 5 --no-roofline 2>>gpurun_out/bench.err | tee gpurun_out/bench_st3.json | cut -c1-330
echo "== pdl0 B2"; SFB_PDL=0 $B --steps 30 --warmup 5 --no-roofline 2>>gpurun_out/bench.err | tee gpurun_out/bench_pdl0.json | cut -c1-330
tail -5 gpurun_out/bench.err
