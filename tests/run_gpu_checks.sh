#!/bin/bash
# Runs every kernel parity check in its own process (a trapped kernel kills its CUDA context),
# each under a timeout, and collects one JSON line per check in gpurun_out/kernel_checks.jsonl.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/kernel_checks.jsonl
: > $OUT
NAMES=$(python - <<'PY'
import sys, os
sys.path.insert(0, "tests")
import kernel_checks
print(" ".join(kernel_checks.CHECKS))
PY
)
for n in $NAMES; do
  timeout 240 python tests/kernel_checks.py $n >> $OUT 2> gpurun_out/check_$n.err || echo "{\"check\": \"$n\", \"exit\": $?}" >> $OUT
  tail -n 3 gpurun_out/check_$n.err | head -c 600 >> $OUT.stderr
done
cat $OUT
