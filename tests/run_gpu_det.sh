#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -n 4 2>&1 | tail -8 > gpurun_out/det_kernels.log; cat gpurun_out/det_kernels.log
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_svd_gpu.py tests/test_reference_pin.py tests/test_clip_gpu.py tests/test_vae_gpu.py -m gpu -x -q -k "tiny or replay or frozen or lora or pin or reference or ln or group" 2>&1 | tail -12 > gpurun_out/det_models.log; cat gpurun_out/det_models.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-extras 2>gpurun_out/det_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['ms_per_step'], d['e2e'].get('ms_per_step'), d['roofline']['achieved'], d['roofline']['time_share_by_entry_point'])"
timeout 300 python tests/graph_breakdown.py 2 2>/dev/null | head -3
