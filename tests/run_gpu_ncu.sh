#!/bin/bash
# Profiling run (round 2): launch list of one eager UNet step (time + DRAM bytes per launch) and
# full-section captures of each kernel family, on kernel-check invocations that launch exactly that kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
KREGEX='regex:gemm_tc|gemm_persist|attention_|gn_|layer_norm|small_linear|conv_in|conv_out|upsample2x|timestep_embed|splitk|im2col|temporal_|row_op|row_softmax|pointwise|add_nchw|copy2d|embed_tokens|clip_pool|patchify'
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$KREGEX" -c 1400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-roofline --no-extras > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log
full() {  # name  kernel-regex  skip  count  command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$rx" -s $skip -c $cnt -o gpurun_out/r02_prof_$name "$@" > gpurun_out/ncu_$name.log 2>&1; tail -1 gpurun_out/ncu_$name.log
}
full gemm_conv64 gemm_tc 0 1 python tests/kernel_checks.py conv_pair_64
full gemm_big gemm_tc 0 1 python tests/kernel_checks.py gemm_pair_big
full gemm_persist gemm_persist 0 1 python tests/kernel_checks.py persist_conv_64
full attn_d40 attention_tc 0 1 python tests/kernel_checks.py attn_d40_4096
full attn_d64 attention_v2 0 1 python tests/kernel_checks.py attn_v2_d64_4096
full gn_fused gn_fused 0 1 python tests/kernel_checks.py group_norm_fused_64x64
full gn_two_pass 'gn_stats|gn_apply' 0 2 python tests/kernel_checks.py group_norm_two_pass_big
full temporal_attn temporal_attention 0 1 python tests/kernel_checks.py temporal_attn_25
full geglu gemm_tc 0 1 python tests/kernel_checks.py geglu
SFB_SHAPES="linear M32768 N320 K320" full shortk_legacy gemm_tc 0 1 python tests/gemm_shapes_bench.py
SFB_SHAPES="geglu M32768 K320" full geglu320_legacy gemm_tc 0 1 python tests/gemm_shapes_bench.py
SFB_SHAPES="geglu M32768 K320" full geglu320_persist gemm_persist 0 1 python tests/gemm_shapes_bench.py
ls -la gpurun_out/*.ncu-rep | head -20
