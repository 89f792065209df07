#!/bin/bash
# Round-2 GPU call 2: new kernels first (under short timeouts), then the whole suite, then A/B benches.
mkdir -p gpurun_out
timeout 180 python tools/attn_check.py > gpurun_out/r2_attn_check.log 2>&1; echo "attn_check rc=$?"; tail -16 gpurun_out/r2_attn_check.log | cut -c1-330
timeout 120 python -m pytest tests/test_bf16.py -q -x -k "stride2 or attention" -p no:cacheprovider > gpurun_out/r2_s2_attn_tests.log 2>&1; echo "s2/attn tests rc=$?"; tail -5 gpurun_out/r2_s2_attn_tests.log | cut -c1-300
timeout 240 python tools/gemm_bench.py > gpurun_out/r2_gemm_bench2.log 2>&1; cat gpurun_out/r2_gemm_bench2.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_call2_gpu_tests.log 2>&1; echo "suite rc=$?"; tail -15 gpurun_out/r2_call2_gpu_tests.log | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c2_default.json 2> gpurun_out/r2_bench_c2_default.err; echo "bench rc=$?"
for f in TFB_ATTN_FUSED TFB_WGRAD_STREAM TFB_CONV_S2_TC; do
  env $f=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c2_${f}_off.json 2> gpurun_out/r2_bench_c2_${f}_off.err
done
TFB_GEMM_BN=128 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c2_bn128.json 2> gpurun_out/r2_bench_c2_bn128.err
for f in gpurun_out/r2_bench_c2_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['roofline']['top5_ms'], d['config']['cuda_graph_error'])
except Exception as e: print('ERR', e); print(open('$f'.replace('.json','.err')).read()[-1500:])
"; done
timeout 400 python tools/bf16_vs_oracle.py > gpurun_out/r2_bf16_vs_oracle.log 2>&1; tail -70 gpurun_out/r2_bf16_vs_oracle.log
