#!/bin/bash
# Round-2 GPU call 4: per-CTA statistics accumulators, wgrad-stream CTA cap, narrow dgrad on tensor cores; suite + smoke; A/B benches.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_bf16.py -q -x -k "statistics or operand_oracle or close_to_fp32 or tensor_core" -p no:cacheprovider > gpurun_out/r2_c4_new_tests.log 2>&1; echo "new tests rc=$?"; grep -E "bf16 mode vs|losses:|passed|failed|Error" gpurun_out/r2_c4_new_tests.log | cut -c1-600
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_c4_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_c4_smoke.log | cut -c1-300
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c4_$name.json 2> gpurun_out/r2_bench_c4_$name.err; }
run default TFB_X=0
run bnstats_off TFB_BN_STATS_FUSED=0
run wgrad_cap32 TFB_WGRAD_MAX_CTAS=32
run wgrad_cap64 TFB_WGRAD_MAX_CTAS=64
run wgrad_cap96 TFB_WGRAD_MAX_CTAS=96
run narrow_off TFB_NARROW_DGRAD_TC=0
for f in gpurun_out/r2_bench_c4_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); r=d['roofline']; print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], r.get('kernel_ms_per_step'), r.get('frac'), r.get('roof_frac'), r.get('large_launches',{}).get('achieved'), d['config']['cuda_graph_error'])
except Exception as e: print('ERR', e); print(open('$f'.replace('.json','.err')).read()[-1500:])
"; done
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_call4_gpu_tests.log 2>&1; echo "suite rc=$?"; tail -8 gpurun_out/r2_call4_gpu_tests.log | cut -c1-300
