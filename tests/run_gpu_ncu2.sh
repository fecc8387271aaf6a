#!/bin/bash
# Profiling run for the folded-GroupNorm path: launch list of one eager UNet step with the default policy
# (time + DRAM bytes per launch), full-section captures of the halo conv and of the statistics kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
KREGEX='regex:gemm_tc|gemm_persist|attention_|gn_|layer_norm|small_linear|conv_in|conv_out|upsample2x|timestep_embed|splitk|im2col|temporal_|row_op|row_softmax|pointwise|add_nchw|copy2d|embed_tokens|clip_pool|patchify'
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KREGEX" -c 1400 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-roofline --no-extras > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KREGEX" -c 1400 --csv --log-file gpurun_out/r02b_launches_b8.csv python bench.py --steps 1 --warmup 3 --batch 8 --no-graph --no-cpu-baseline --no-roofline --no-extras > gpurun_out/ncu_bench_b8.log 2>&1; tail -2 gpurun_out/ncu_bench_b8.log | cut -c1-300
full() {  # name  kernel-regex  skip  count  command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $cnt -o gpurun_out/r02b_prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1; tail -1 gpurun_out/ncu_$name.log
}
SFB_ONLY=folded SFB_SHAPES="n8 64x64 960->320" full conv_gn_960 'gemm_tc|gn_stats_ab' 0 2 python tests/conv_gn_bench.py
SFB_ONLY=folded SFB_SHAPES="n8 64x64 320->320" full conv_gn_320 'gemm_tc|gn_stats_ab' 0 2 python tests/conv_gn_bench.py
SFB_ONLY=separate SFB_SHAPES="n8 64x64 960->320" full conv_9tap_960 'gemm_tc' 0 1 python tests/conv_gn_bench.py
ls -la gpurun_out/*.ncu-rep | head
