#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -n 4 --timeout 300 2>&1 | tail -40 > gpurun_out/pytest_kernels.log; tail -5 gpurun_out/pytest_kernels.log
timeout 1200 python -m pytest tests/test_vae_gpu.py "tests/test_unet_gpu.py::test_in_place_parameter_update_is_picked_up_like_the_reference_lora_contract" tests/test_unet_gpu.py::test_tiny_unet_vs_oracle -m gpu -q -n 3 --timeout 900 2>&1 | tail -40 > gpurun_out/pytest_vae.log; tail -12 gpurun_out/pytest_vae.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
cut -c1-600 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
