#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_call12_gpu_tests.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/r2_call12_gpu_tests.log | cut -c1-300
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c12_$name.json 2> gpurun_out/r2_bench_c12_$name.err; }
run default TFB_X=0
run cap48 TFB_WGRAD_MAX_CTAS=48
run cap80 TFB_WGRAD_MAX_CTAS=80
run overlap_off TFB_OVERLAP_ADAMW=0
for f in gpurun_out/r2_bench_c12_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); r=d['roofline']; print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], r.get('kernel_ms_per_step'), r.get('frac'), r.get('roof_frac'), r.get('large_launches',{}).get('achieved'), d['config']['cuda_graph_error'])
except Exception as e: print('ERR', e); print(open('$f'.replace('.json','.err')).read()[-1200:])
"; done
