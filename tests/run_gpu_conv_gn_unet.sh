#!/bin/bash
# Folded-GroupNorm conv at the UNet level, same box: bench lines with SFB_CONV_GN = 0 (GroupNorm kernels +
# 9-tap convs) and auto (the default policy) at 8 latents per GPU (64^2, 128^2) and SDXL 8 x 128^2 bf16.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/conv_gn_unet_bench.jsonl
run() {  # tag, bench args...
  local tag=$1; shift
  for mode in 0 auto; do
    SFB_CONV_GN=$mode timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras "$@" 2> gpurun_out/conv_gn_unet_${tag}_$mode.err | python -c "
import json,sys
for line in sys.stdin:
    try: d=json.loads(line)
    except Exception: continue
    r=d.get('roofline') or {}
    print(json.dumps({'workload':'$tag','conv_gn':'$mode','ms_per_step':d.get('ms_per_step'),'value':d.get('value'),'gemm_family_tflops':r.get('achieved'),'time_share':r.get('time_share_by_entry_point'),'clocks':d.get('clocks')}))
" >> gpurun_out/conv_gn_unet_bench.jsonl
    tail -n 1 gpurun_out/conv_gn_unet_${tag}_$mode.err | head -c 300
  done
}
run sd15_b8_64 --batch 8
run sd15_b8_128 --batch 8 --size 128
run sdxl_b8_128_bf16 --model sdxl --batch 8 --size 128 --dtype bf16
cat gpurun_out/conv_gn_unet_bench.jsonl
