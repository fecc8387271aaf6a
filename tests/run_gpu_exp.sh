#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python tests/kernel_checks.py > gpurun_out/kernel_checks.jsonl 2> gpurun_out/kernel_checks.err
echo "checks rc=$?" >> gpurun_out/kernel_checks.jsonl
grep -c '"pass": true' gpurun_out/kernel_checks.jsonl; grep -v '"pass": true' gpurun_out/kernel_checks.jsonl | cut -c1-300
timeout 300 python tests/attn_bench.py > gpurun_out/attn_bench.jsonl 2>gpurun_out/attn_bench.err; cat gpurun_out/attn_bench.jsonl
timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --dump-ops gpurun_out/ops_b2.jsonl > gpurun_out/bench.json 2>gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json
timeout 300 python bench.py --steps 20 --warmup 3 --batch 16 --no-cpu-baseline --dump-ops gpurun_out/ops_b16.jsonl > gpurun_out/bench_b16.json 2>>gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_b16.json
timeout 300 python tests/graph_breakdown.py 2 > gpurun_out/breakdown_b2.jsonl 2>gpurun_out/breakdown.err; head -8 gpurun_out/breakdown_b2.jsonl; tail -2 gpurun_out/breakdown_b2.jsonl
tail -3 gpurun_out/bench.err
KREGEX='regex:gemm_tc|attention_tc|gn_|layer_norm|small_linear|conv_in|conv_out|upsample2x|timestep_embed|splitk|im2col'
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KREGEX" -c 1400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 700 -c 4 -o gpurun_out/prof_gemm_r01 python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-roofline > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
