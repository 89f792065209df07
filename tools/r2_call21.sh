#!/bin/bash
# call 21: one-launch sliced BatchNorm backward + BatchNorm/squeeze-excite as one node: tests, A/B bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops.py tests/test_bf16.py tests/test_trainer.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_call21_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_call21_tests.log | cut -c1-300
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c21_$name.json 2> gpurun_out/r2_bench_c21_$name.err; }
run default TFB_X=0
run sliced_off TFB_BN_BWD_SLICED=0
run bnse_off TFB_BN_SE_FUSED=0
run both_off TFB_BN_SE_FUSED=0 TFB_BN_BWD_SLICED=0
for f in gpurun_out/r2_bench_c21_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); r=d['roofline']; print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], r.get('kernel_ms_per_step'), r.get('frac'), r.get('roof_frac'), d['config']['cuda_graph_error'])
except Exception as e: print('ERR', e); print(open('$f'.replace('.json','.err')).read()[-1200:])
"; done
