#!/bin/bash
cd "$(dirname "$0")/.."
bash tests/run_gpu_full.sh
timeout 300 python tests/conv_gn_bench.py > gpurun_out/conv_gn_bench.jsonl 2> gpurun_out/conv_gn_bench.err; cat gpurun_out/conv_gn_bench.jsonl
