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
timeout 300 python tests/kernel_checks.py $NAMES 2> gpurun_out/kernel_checks.err | tee gpurun_out/kernel_checks_attn.jsonl | cut -c1-130
tail -3 gpurun_out/kernel_checks.err
timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -s -k "sdxl" 2>&1 | grep -v "^$" | tail -5
echo "== SDXL B=8 128 bf16"; timeout 600 python bench.py --no-cpu-baseline --model sdxl --batch 8 --size 128 --dtype bf16 --steps 10 --warmup 3 2>>gpurun_out/bench.err | tee gpurun_out/bench_sdxl_b8.json | cut -c1-1800
tail -3 gpurun_out/bench.err
