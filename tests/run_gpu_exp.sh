#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NONPAIR=$(python - <<'PY'
import sys
sys.path.insert(0, "tests")
import kernel_checks
print(" ".join(n for n in kernel_checks.CHECKS if "pair" not in n))
PY
)
PAIR=$(python - <<'PY'
import sys
sys.path.insert(0, "tests")
import kernel_checks
print(" ".join(n for n in kernel_checks.CHECKS if "pair" in n))
PY
)
timeout 900 python tests/kernel_checks.py $NONPAIR > gpurun_out/kernel_checks.jsonl 2> gpurun_out/kernel_checks.err
echo "checks rc=$?" >> gpurun_out/kernel_checks.jsonl
grep -c '"pass": true' gpurun_out/kernel_checks.jsonl; grep -v '"pass": true' gpurun_out/kernel_checks.jsonl | cut -c1-300
: > gpurun_out/pair_checks.jsonl
for n in $PAIR; do
  timeout 90 python tests/kernel_checks.py $n >> gpurun_out/pair_checks.jsonl 2> gpurun_out/pair_$n.err || echo "{\"check\": \"$n\", \"exit\": $?}" >> gpurun_out/pair_checks.jsonl
done
cat gpurun_out/pair_checks.jsonl | cut -c1-200
timeout 300 python tests/gemm_latency.py > gpurun_out/gemm_latency.jsonl 2>gpurun_out/gemm_latency.err; cat gpurun_out/gemm_latency.jsonl | cut -c1-420
timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -5
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --dump-ops gpurun_out/ops_b2.jsonl > gpurun_out/bench.json 2>gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json
timeout 300 python bench.py --steps 20 --warmup 3 --batch 16 --no-cpu-baseline --dump-ops gpurun_out/ops_b16.jsonl > gpurun_out/bench_b16.json 2>>gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_b16.json
timeout 300 python tests/graph_breakdown.py 2 > gpurun_out/breakdown_b2.jsonl 2>gpurun_out/breakdown.err; head -12 gpurun_out/breakdown_b2.jsonl; tail -2 gpurun_out/breakdown_b2.jsonl
