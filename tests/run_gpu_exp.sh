#!/bin/bash
# Kernel checks with clusters on, then an experiment matrix (cluster x PDL x GEMM stage policy) on
# the B = 2 graph step time, then an ncu capture of the attention kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python tests/kernel_checks.py > gpurun_out/kernel_checks.jsonl 2> gpurun_out/kernel_checks.err
echo "checks rc=$?" >> gpurun_out/kernel_checks.jsonl
grep -c '"pass": true' gpurun_out/kernel_checks.jsonl; grep -v '"pass": true' gpurun_out/kernel_checks.jsonl | cut -c1-300
: > gpurun_out/exp.jsonl
for cl in 1 0; do for pdl in 1 0; do for st in 0 3; do
  SFB_CLUSTER=$cl SFB_PDL=$pdl SFB_GEMM_STAGES=$st timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-roofline 2>gpurun_out/exp.err | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(json.dumps({'cluster': $cl, 'pdl': $pdl, 'stages': $st, 'ms_per_step': round(d['ms_per_step'],3), 'e2e_ms': round(d['e2e']['ms_per_step'],3), 'launches': d['kernel_launches_per_step']}))" >> gpurun_out/exp.jsonl
done; done; done
SFB_CLUSTER=1 timeout 300 python bench.py --steps 20 --warmup 3 --batch 16 --no-cpu-baseline --no-roofline 2>>gpurun_out/exp.err | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(json.dumps({'b16_cluster': 1, 'ms_per_step': d['ms_per_step']}))" >> gpurun_out/exp.jsonl
SFB_CLUSTER=0 timeout 300 python bench.py --steps 20 --warmup 3 --batch 16 --no-cpu-baseline --no-roofline 2>>gpurun_out/exp.err | python -c "
import sys, json
d=json.loads(sys.stdin.read()); print(json.dumps({'b16_cluster': 0, 'ms_per_step': d['ms_per_step']}))" >> gpurun_out/exp.jsonl
cat gpurun_out/exp.jsonl; tail -3 gpurun_out/exp.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 4 -c 2 -o gpurun_out/prof_attention python tests/attn_bench.py > gpurun_out/ncu_attn.log 2>&1
tail -2 gpurun_out/ncu_attn.log
