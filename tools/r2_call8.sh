#!/bin/bash
# Round-2 GPU call 8: cross-stream sidecar race fix, vectorised upsample kernels, split-K chooser under the CTA cap.
mkdir -p gpurun_out
timeout 300 python tools/graph_vs_eager.py default > gpurun_out/r2_graph_vs_eager2.log 2>&1; echo "rc=$?"; grep -E "^default|graph step3" gpurun_out/r2_graph_vs_eager2.log | cut -c1-400
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider -s > gpurun_out/r2_call8_gpu_tests.log 2>&1; echo "suite rc=$?"; grep -E "eager vs eager|bf16 mode vs|passed|failed" gpurun_out/r2_call8_gpu_tests.log | cut -c1-400 | tail -8
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_c8_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_c8_smoke.log | cut -c1-300
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c8_$name.json 2> gpurun_out/r2_bench_c8_$name.err; }
run default TFB_X=0
run autosplit TFB_WGRAD_AUTO_SPLIT=1
run autosplit_cap64 TFB_WGRAD_AUTO_SPLIT=1 TFB_WGRAD_MAX_CTAS=64
run decoder_one_stream TFB_DECODER_STREAMS=0
for f in gpurun_out/r2_bench_c8_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); r=d['roofline']; print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], r.get('kernel_ms_per_step'), r.get('frac'), r.get('roof_frac'), r.get('large_launches',{}).get('achieved'), d['config']['cuda_graph_error'])
except Exception as e: print('ERR', e); print(open('$f'.replace('.json','.err')).read()[-1200:])
"; done
