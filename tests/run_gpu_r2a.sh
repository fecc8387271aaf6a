#!/bin/bash
# Round-2 validation run: kernel parity checks (4 worker processes: a trapped kernel only poisons its
# own worker's CUDA context), whole-UNet / reference-pin tests, smoke, the bench line with extras.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -n 4 --timeout 300 2>&1 | tail -40 > gpurun_out/pytest_kernels.log; tail -6 gpurun_out/pytest_kernels.log
timeout 1500 python -m pytest tests/test_reference_pin.py tests/test_unet_gpu.py -m gpu -q -n 3 --timeout 900 --durations=8 2>&1 | tail -60 > gpurun_out/pytest_unet.log; tail -14 gpurun_out/pytest_unet.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
cut -c1-1200 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
