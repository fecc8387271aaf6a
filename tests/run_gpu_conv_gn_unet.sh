#!/bin/bash
# Folded-GroupNorm conv at the UNet level: whole-model parity tests with the default policy, then
# A/B bench lines (SFB_CONV_GN = 0 / auto / 1) at the headline config and at 8 latents per GPU.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_vae_gpu.py -x -q -m gpu --durations=8 ) 2>&1 | grep -v "CUDAEvent.h" | tail -25 > gpurun_out/conv_gn_unet_pytest.log
cat gpurun_out/conv_gn_unet_pytest.log
: > gpurun_out/conv_gn_unet_bench.jsonl
for mode in 0 auto 1; do
  for batch in 2 8; do
    SFB_CONV_GN=$mode timeout 300 python bench.py --steps 20 --warmup 5 --batch $batch --no-cpu-baseline --no-extras 2> gpurun_out/conv_gn_unet_bench_${mode}_$batch.err | python -c "
import json,sys
for line in sys.stdin:
    try: d=json.loads(line)
    except Exception: continue
    print(json.dumps({'conv_gn':'$mode','batch':$batch,'ms_per_step':d.get('ms_per_step'),'value':d.get('value'),'roofline':d.get('roofline'),'gpu_launches':d.get('gpu_launches'),'e2e':d.get('e2e')}))
" >> gpurun_out/conv_gn_unet_bench.jsonl
    tail -n 2 gpurun_out/conv_gn_unet_bench_${mode}_$batch.err | head -c 400
  done
done
cat gpurun_out/conv_gn_unet_bench.jsonl
