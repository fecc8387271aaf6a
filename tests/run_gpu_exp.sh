#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NAMES=$(python - <<'PY'
import re
src = open("tests/kernel_checks.py").read()
names = re.findall(r'^    "([a-z0-9_]+)": \(', src, re.M)
print(" ".join(n for n in names if re.search(r"^conv|^upconv|^gn_finish|^gemm_small|^gemm_pair$", n)))
PY
)
timeout 600 python tests/kernel_checks.py $NAMES > gpurun_out/kernel_checks_conv.jsonl 2> gpurun_out/kernel_checks.err
echo "checks rc=$?"; grep -c '"pass": true' gpurun_out/kernel_checks_conv.jsonl; grep -v '"pass": true' gpurun_out/kernel_checks_conv.jsonl | cut -c1-400
tail -5 gpurun_out/kernel_checks.err
timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -k "not sdxl_unet_full and not 128_latent" 2>&1 | tail -5
B="timeout 300 python bench.py --no-cpu-baseline --no-roofline"
echo "== default B2"; $B --steps 30 --warmup 5 2>>gpurun_out/bench.err | tee gpurun_out/bench_patch.json | cut -c1-330
echo "== SD15 96x96 B=2 (768^2)"; $B --size 96 --steps 20 --warmup 3 2>>gpurun_out/bench.err | tee gpurun_out/bench_sd15_96.json | cut -c1-420
tail -5 gpurun_out/bench.err
