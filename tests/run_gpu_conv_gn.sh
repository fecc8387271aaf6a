#!/bin/bash
# Folded-GroupNorm halo conv: parity checks (own process each, a trapped kernel kills its context), then per-shape timings against GroupNorm kernel + 9-tap conv.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OUT=gpurun_out/conv_gn_checks.jsonl
: > $OUT
for n in conv_gn_16_one_block conv_gn_64 conv_gn_32_640 conv_gn_concat_pitch conv_gn_24_ragged_rows conv_gn_40x24_ragged_n conv_gn_mean50 conv_gn_nosilu conv_gn_b8 conv_gn_bf16; do
  timeout 120 python tests/kernel_checks.py $n >> $OUT 2> gpurun_out/conv_gn_$n.err || { echo "{\"check\": \"$n\", \"exit\": $?}" >> $OUT; tail -n 5 gpurun_out/conv_gn_$n.err | head -c 800 >> $OUT; }
done
cat $OUT
if grep -q '"check": "conv_gn_64".*"pass": true' $OUT; then
  timeout 300 python tests/conv_gn_bench.py > gpurun_out/conv_gn_bench.jsonl 2> gpurun_out/conv_gn_bench.err
  cat gpurun_out/conv_gn_bench.jsonl; tail -n 5 gpurun_out/conv_gn_bench.err
fi
