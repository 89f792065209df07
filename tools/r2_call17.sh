#!/bin/bash
# call 17: the tensor-core parity mode (bf16x3): per-op tests, whole-model parity test in both exact modes, smoke, bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bf16x3.py tests/test_model.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_call17_tests.log 2>&1; echo "tests rc=$?"; grep -n "^\[bf16x3\]\|^\[simt\]\|passed\|failed" gpurun_out/r2_call17_tests.log | cut -c1-200 | tail -60
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_call17_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2_call17_smoke.log | cut -c1-300
timeout 600 python bench.py --gemm bf16x3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c17_bf16x3.json 2> gpurun_out/r2_bench_c17_bf16x3.err; echo "bench rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_c17_bf16x3.json') if l.startswith('{')][-1]); r=d['roofline']; print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['dtype'], r.get('kernel'), r.get('kernel_ms_per_step'), d['config']['cuda_graph_error'])
" || tail -20 gpurun_out/r2_bench_c17_bf16x3.err
