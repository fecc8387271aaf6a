#!/bin/bash
# One GPU-box session: isolated kernel checks, parity tests, smoke, bench, ncu launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
if [ "$RUN_MICRO" == "1" ]; then
  timeout 120 tests/microbench/softmax_pipes > gpurun_out/softmax_pipes.jsonl 2>&1
  SFB_ATTN_EXP16=0 timeout 300 python tests/attn_bench.py > gpurun_out/attn_bench.jsonl 2>gpurun_out/attn_bench.err
  SFB_ATTN_EXP16=1 timeout 300 python tests/attn_bench.py >> gpurun_out/attn_bench.jsonl 2>>gpurun_out/attn_bench.err
  cat gpurun_out/softmax_pipes.jsonl gpurun_out/attn_bench.jsonl
fi
if [ "$RUN_CHECKS" == "1" ]; then
  timeout 600 python tests/kernel_checks.py > gpurun_out/kernel_checks.jsonl 2> gpurun_out/kernel_checks.err
  echo "checks rc=$?" >> gpurun_out/kernel_checks.jsonl
  grep -c '"pass": true' gpurun_out/kernel_checks.jsonl; grep -v '"pass": true' gpurun_out/kernel_checks.jsonl | cut -c1-300
fi
if [ "$SKIP_PYTEST" != "1" ]; then
timeout 1500 python -m pytest tests -m gpu -q -s ${PYTEST_ARGS} 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
fi
timeout 900 python bench.py --steps 50 --warmup 5 --dump-ops gpurun_out/ops_b2.jsonl > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python bench.py --steps 20 --warmup 3 --batch 16 --no-cpu-baseline --dump-ops gpurun_out/ops_b16.jsonl > gpurun_out/bench_b16.json 2>> gpurun_out/bench.err
KREGEX='regex:gemm_tc|attention_tc|gn_|layer_norm|small_linear|conv_in|conv_out|upsample2x|timestep_embed|splitk'
if [ "$SKIP_NCU" != "1" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 1200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench.log 2>&1
fi
if [ "$NCU_FULL" != "" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$NCU_FULL -s ${NCU_SKIP:-60} -c ${NCU_COUNT:-3} -o gpurun_out/prof_$NCU_FULL python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-roofline > gpurun_out/ncu_full.log 2>&1
fi
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -3; cat gpurun_out/bench.json | cut -c1-900; tail -3 gpurun_out/bench.err; cat gpurun_out/bench_b16.json | cut -c1-300
