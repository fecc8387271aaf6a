#!/bin/bash
# End-of-round evidence run on one B200: full GPU test suite, smoke, bench lines, graph breakdown,
# ncu launch list (+ DRAM bytes) and full-section captures of the GEMM and attention kernels.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 50 --warmup 10 --dump-ops gpurun_out/ops_b2.jsonl > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
cut -c1-3000 gpurun_out/bench.json
timeout 600 python bench.py --steps 20 --warmup 5 --batch 16 --no-cpu-baseline > gpurun_out/bench_b16.json 2>> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_b16.json
if [ -f stable-fast_b200/sfast_b200/libsfb200_prev.so ]; then  # optional A/B against a previous build
  SFB_LIB_PATH=$PWD/stable-fast_b200/sfast_b200/libsfb200_prev.so timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-roofline 2>>gpurun_out/bench.err | tee gpurun_out/bench_prev_lib.json | cut -c1-330
  timeout 300 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --no-roofline 2>>gpurun_out/bench.err | tee gpurun_out/bench_new_lib.json | cut -c1-330
fi
timeout 300 python tests/graph_breakdown.py 2 > gpurun_out/breakdown_b2.jsonl 2>gpurun_out/breakdown.err; head -6 gpurun_out/breakdown_b2.jsonl; tail -2 gpurun_out/breakdown_b2.jsonl
KREGEX='regex:gemm_tc|attention_tc|gn_|layer_norm|small_linear|conv_in|conv_out|upsample2x|timestep_embed|splitk|im2col'
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KREGEX" -c 1300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench.log 2>&1
if [ "$NCU_FULL" == "1" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 700 -c 6 -o gpurun_out/prof_gemm_r01 python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-roofline > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 96 -c 3 -o gpurun_out/prof_attention_r01 python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-roofline > gpurun_out/ncu_full_attn.log 2>&1; tail -1 gpurun_out/ncu_full_attn.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_fused -s 180 -c 3 -o gpurun_out/prof_gn_r01 python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-roofline > gpurun_out/ncu_full_gn.log 2>&1; tail -1 gpurun_out/ncu_full_gn.log
fi
tail -3 gpurun_out/bench.err
