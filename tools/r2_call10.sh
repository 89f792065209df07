#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" TFB_CAPTURE_DEBUG=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c10_$name.json 2> gpurun_out/r2_bench_c10_$name.err; }
run default TFB_X=0
run autosplit TFB_WGRAD_AUTO_SPLIT=1
run autosplit_cap64 TFB_WGRAD_AUTO_SPLIT=1 TFB_WGRAD_MAX_CTAS=64
run decoder_one_stream TFB_DECODER_STREAMS=0
for f in gpurun_out/r2_bench_c10_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); r=d['roofline']; print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], r.get('kernel_ms_per_step'), r.get('frac'), r.get('roof_frac'), r.get('large_launches',{}).get('achieved'), d['config']['cuda_graph_error'])
except Exception as e: print('ERR', e); print(open('$f'.replace('.json','.err')).read()[-1200:])
"; done
timeout 300 python -m pytest tests/test_trainer.py -m gpu -q -p no:cacheprovider -s 2>&1 | grep -E "eager vs eager|passed|failed" | cut -c1-300
