#!/bin/bash
cd "$(dirname "$0")/.."
bash tests/run_gpu_n1a.sh
if grep -q '"pass": false\|"exit"' gpurun_out/n1a_checks.jsonl; then echo "checks failed: skipping the UNet-level run"; exit 1; fi
bash tests/run_gpu_n1b.sh
