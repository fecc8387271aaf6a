#!/bin/bash
# What the driver runs at round end, in its form: the whole GPU suite in one process, smoke(), bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
( time timeout 1800 python -m pytest tests/ -x -q -m gpu --durations=12 ) 2>&1 | grep -v "CUDAEvent.h" > gpurun_out/r02_pytest_gpu_full.log; tail -40 gpurun_out/r02_pytest_gpu_full.log > gpurun_out/r02_pytest_gpu.log; grep -n "Error\|error:" gpurun_out/r02_pytest_gpu_full.log | head -20; tail -22 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke.log; tail -4 gpurun_out/r02_smoke.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
