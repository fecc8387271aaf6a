#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NAMES=$(python - <<'PY'
import re
src = open("tests/kernel_checks.py").read()
names = re.findall(r'^    "([a-z0-9_]+)": \(', src, re.M)
print(" ".join(n for n in names if re.search(r"^attn", n)))
PY
)
echo "== attention checks, v2"
SFB_ATTN_V2=1 timeout 300 python tests/kernel_checks.py $NAMES 2> gpurun_out/kernel_checks.err | tee gpurun_out/kernel_checks_attn_v2.jsonl | cut -c1-160
tail -3 gpurun_out/kernel_checks.err
echo "== attn bench v2"; SFB_ATTN_V2=1 timeout 200 python tests/attn_bench.py 2>gpurun_out/attn_bench.err | tee gpurun_out/attn_bench_v2.jsonl
echo "== attn bench v1"; timeout 200 python tests/attn_bench.py 2>>gpurun_out/attn_bench.err | tee gpurun_out/attn_bench_v1.jsonl
B="timeout 300 python bench.py --no-cpu-baseline --no-roofline --steps 30 --warmup 5"
echo "== v2 B2"; SFB_ATTN_V2=1 $B 2>>gpurun_out/bench.err | tee gpurun_out/bench_av2.json | cut -c1-330
echo "== v1 B2"; $B 2>>gpurun_out/bench.err | tee gpurun_out/bench_av1.json | cut -c1-330
echo "== v2 B16"; SFB_ATTN_V2=1 $B --batch 16 --steps 20 2>>gpurun_out/bench.err | tee gpurun_out/bench_b16_av2.json | cut -c1-330
echo "== v1 B16"; $B --batch 16 --steps 20 2>>gpurun_out/bench.err | tee gpurun_out/bench_b16_av1.json | cut -c1-330
SFB_ATTN_V2=1 timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -k "sd15_unet_vs_oracle_full_size or tiny_unet_vs_oracle or bf16_tiny or sdxl_tiny or rectangular" 2>&1 | tail -3
tail -3 gpurun_out/bench.err
