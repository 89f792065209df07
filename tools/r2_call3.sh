#!/bin/bash
# Round-2 GPU call 3: BN statistics out of the conv / GEMM epilogues, bf16-operand oracle parity, new smoke(), other BASELINE configs.
mkdir -p gpurun_out
timeout 240 python tools/attn_check.py > gpurun_out/r2_attn_check2.log 2>&1; echo "attn_check rc=$?"; grep -E "B10|dropout|bf16 inputs" gpurun_out/r2_attn_check2.log | cut -c1-330
timeout 300 python -m pytest tests/test_bf16.py -q -x -k "statistics or stride2 or operand_oracle or sidecars" -p no:cacheprovider > gpurun_out/r2_c3_new_tests.log 2>&1; echo "new tests rc=$?"; grep -E "bf16 mode vs|losses:|run-to-run|passed|failed|Error" gpurun_out/r2_c3_new_tests.log | cut -c1-600
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_c3_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2_c3_smoke.log | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c3_default.json 2> gpurun_out/r2_bench_c3_default.err; echo "bench rc=$?"
TFB_WGRAD_AUTO_SPLIT=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c3_wgradsplit_off.json 2> gpurun_out/r2_bench_c3_wgradsplit_off.err
TFB_BN_STATS_FUSED=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c3_bnstats_off.json 2> gpurun_out/r2_bench_c3_bnstats_off.err
for c in 3 4 5; do
  timeout 300 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c3_config$c.json 2> gpurun_out/r2_bench_c3_config$c.err
done
for f in gpurun_out/r2_bench_c3_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], {k: v for k, v in d['roofline'].items() if k not in ('top5_ms',)}, d['config']['cuda_graph_error'])
except Exception as e: print('ERR', e); print(open('$f'.replace('.json','.err')).read()[-1500:])
"; done
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_call3_gpu_tests.log 2>&1; echo "suite rc=$?"; tail -8 gpurun_out/r2_call3_gpu_tests.log | cut -c1-300
timeout 400 python tools/bf16_vs_oracle.py > gpurun_out/r2_bf16_vs_oracle.log 2>&1; grep -A40 "bf16 vs the bf16-operand" gpurun_out/r2_bf16_vs_oracle.log | cut -c1-200
