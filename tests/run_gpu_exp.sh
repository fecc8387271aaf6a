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
timeout 600 python tests/kernel_checks.py $NAMES > gpurun_out/kernel_checks_attn.jsonl 2> gpurun_out/kernel_checks.err
echo "checks rc=$?"; cat gpurun_out/kernel_checks_attn.jsonl | cut -c1-200
echo "== attn bench new"; timeout 300 python tests/attn_bench.py 2>gpurun_out/attn_bench.err | tee gpurun_out/attn_bench_new.jsonl
echo "== attn bench prev"; SFB_LIB_PATH=$PWD/stable-fast_b200/sfast_b200/libsfb200_prev.so timeout 300 python tests/attn_bench.py 2>>gpurun_out/attn_bench.err | tee gpurun_out/attn_bench_prev.jsonl
B="timeout 300 python bench.py --no-cpu-baseline --no-roofline --steps 30 --warmup 5"
echo "== new (lazy max)"; $B 2>>gpurun_out/bench.err | tee gpurun_out/bench_lazy1.json | cut -c1-330
echo "== prev lib"; SFB_LIB_PATH=$PWD/stable-fast_b200/sfast_b200/libsfb200_prev.so $B 2>>gpurun_out/bench.err | tee gpurun_out/bench_lazy0.json | cut -c1-330
echo "== new B16"; $B --batch 16 --steps 20 2>>gpurun_out/bench.err | tee gpurun_out/bench_b16_lazy1.json | cut -c1-330
timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -k "sd15_unet_vs_oracle_full_size or tiny_unet_vs_oracle or bf16_tiny or sdxl_tiny" 2>&1 | tail -3
tail -3 gpurun_out/bench.err
