#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NAMES=$(python - <<'PY'
import re, sys
src = open("tests/kernel_checks.py").read()
names = re.findall(r'^    "([a-z0-9_]+)": \(', src, re.M)
print(" ".join(n for n in names if re.search(r"group_norm|gn_finish", n)))
PY
)
timeout 600 python tests/kernel_checks.py $NAMES > gpurun_out/kernel_checks_gn.jsonl 2> gpurun_out/kernel_checks.err
echo "checks rc=$?"; grep -c '"pass": true' gpurun_out/kernel_checks_gn.jsonl; grep -v '"pass": true' gpurun_out/kernel_checks_gn.jsonl | cut -c1-400
tail -5 gpurun_out/kernel_checks.err
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-roofline"
echo "== default (gn per-group on)"; $B 2>>gpurun_out/bench.err | tee gpurun_out/bench_gg1.json | cut -c1-330
echo "== gn per-group off"; SFB_GN_GROUP=0 $B 2>>gpurun_out/bench.err | tee gpurun_out/bench_gg0.json | cut -c1-330
echo "== B16 default"; $B --batch 16 --steps 20 2>>gpurun_out/bench.err | tee gpurun_out/bench_b16.json | cut -c1-330
echo "== B16 off"; SFB_GN_GROUP=0 $B --batch 16 --steps 20 2>>gpurun_out/bench.err | tee gpurun_out/bench_b16_gg0.json | cut -c1-330
tail -5 gpurun_out/bench.err
timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q -x -k "not sdxl_unet_full and not 128_latent" 2>&1 | tail -4
