#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NAMES=$(python - <<'PY'
import re
src = open("tests/kernel_checks.py").read()
names = re.findall(r'^    "([a-z0-9_]+)": \(', src, re.M)
print(" ".join(n for n in names if re.search(r"^gemm|^conv|^geglu|^ln_fold|^upconv|^kv_scatter|^qkv", n)))
PY
)
timeout 600 python tests/kernel_checks.py $NAMES > gpurun_out/kernel_checks_gemm.jsonl 2> gpurun_out/kernel_checks.err
echo "checks rc=$?"; grep -c '"pass": true' gpurun_out/kernel_checks_gemm.jsonl; grep -v '"pass": true' gpurun_out/kernel_checks_gemm.jsonl | cut -c1-400
tail -5 gpurun_out/kernel_checks.err
B="timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-roofline"
echo "== default (pipelined epilogue)"; $B 2>>gpurun_out/bench.err | tee gpurun_out/bench_pipe1.json | cut -c1-330
echo "== no pipeline"; SFB_LIB_PATH=$PWD/stable-fast_b200/sfast_b200/libsfb200_nopipe.so $B 2>>gpurun_out/bench.err | tee gpurun_out/bench_pipe0.json | cut -c1-330
echo "== default again"; $B 2>>gpurun_out/bench.err | tee gpurun_out/bench_pipe1b.json | cut -c1-330
tail -5 gpurun_out/bench.err
timeout 300 python tests/gemm_latency.py 2>gpurun_out/gemm_latency.err | head -3 | cut -c1-520
