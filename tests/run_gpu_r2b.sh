#!/bin/bash
# Round-2 validation run B: kernel checks (4 workers), UNet / SVD / reference-pin tests, graph breakdowns.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -n 4 --timeout 300 2>&1 | tail -60 > gpurun_out/pytest_kernels.log; tail -8 gpurun_out/pytest_kernels.log
timeout 1800 python -m pytest tests/test_svd_gpu.py tests/test_reference_pin.py tests/test_unet_gpu.py -m gpu -q -n 3 --timeout 1200 --durations=8 2>&1 | tail -80 > gpurun_out/pytest_unet.log; tail -16 gpurun_out/pytest_unet.log
timeout 300 python tests/graph_breakdown.py 2 > gpurun_out/breakdown_b2.jsonl 2>gpurun_out/breakdown.err; head -3 gpurun_out/breakdown_b2.jsonl
timeout 300 python tests/graph_breakdown.py 8 > gpurun_out/breakdown_b8.jsonl 2>>gpurun_out/breakdown.err; head -30 gpurun_out/breakdown_b8.jsonl
