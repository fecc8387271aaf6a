#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PAIR=$(python - <<'PY'
import sys
sys.path.insert(0, "tests")
import kernel_checks
print(" ".join(n for n in kernel_checks.CHECKS if "pair" in n))
PY
)
: > gpurun_out/pair_checks.jsonl
for n in $PAIR; do
  timeout 90 python tests/kernel_checks.py $n >> gpurun_out/pair_checks.jsonl 2> gpurun_out/pair_$n.err || echo "{\"check\": \"$n\", \"exit\": $?}" >> gpurun_out/pair_checks.jsonl
done
cat gpurun_out/pair_checks.jsonl | cut -c1-420
SFB_CTA_PAIR=1 timeout 300 python tests/gemm_latency.py > gpurun_out/gemm_latency_pair.jsonl 2>gpurun_out/gemm_latency_pair.err; cat gpurun_out/gemm_latency_pair.jsonl | cut -c1-420; tail -2 gpurun_out/gemm_latency_pair.err
SFB_CTA_PAIR=1 timeout 300 python bench.py --steps 20 --warmup 3 --batch 16 --no-cpu-baseline --dump-ops gpurun_out/ops_b16_pair.jsonl > gpurun_out/bench_b16_pair.json 2>gpurun_out/bench_pair.err; cut -c1-300 gpurun_out/bench_b16_pair.json; tail -2 gpurun_out/bench_pair.err
SFB_CTA_PAIR=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_b2_pair.json 2>>gpurun_out/bench_pair.err; cut -c1-300 gpurun_out/bench_b2_pair.json
