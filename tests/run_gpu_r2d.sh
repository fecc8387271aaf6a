#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q -n 4 --timeout 300 2>&1 | tail -30 > gpurun_out/pytest_kernels.log; tail -4 gpurun_out/pytest_kernels.log
timeout 1200 python -m pytest tests/test_vae_gpu.py -m gpu -q --timeout 900 -s 2>&1 | grep -v "^$" | tail -40 > gpurun_out/pytest_vae.log; tail -14 gpurun_out/pytest_vae.log
timeout 300 python tests/graph_breakdown.py 2 2>gpurun_out/breakdown.err | head -4 > gpurun_out/breakdown_b2_coop.jsonl; cat gpurun_out/breakdown_b2_coop.jsonl
SFB_GN_COOP=0 timeout 300 python tests/graph_breakdown.py 2 2>>gpurun_out/breakdown.err | head -4 > gpurun_out/breakdown_b2_nocoop.jsonl; cat gpurun_out/breakdown_b2_nocoop.jsonl
SFB_PDL=0 timeout 300 python tests/graph_breakdown.py 2 2>>gpurun_out/breakdown.err | head -4 > gpurun_out/breakdown_b2_nopdl.jsonl; cat gpurun_out/breakdown_b2_nopdl.jsonl
rm -f gpurun_out/attn_bench.jsonl; for lib in libsfb200_m00.so libsfb200_m08.so libsfb200.so libsfb200_m92.so; do SFB_LIB_PATH=$PWD/stable-fast_b200/sfast_b200/$lib timeout 300 python tests/attn_bench.py >> gpurun_out/attn_bench.jsonl 2>> gpurun_out/attn_bench.err; done; cat gpurun_out/attn_bench.jsonl
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
