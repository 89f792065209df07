#!/bin/bash
# call 19: parity modes bf16x6 / bf16x3 after the reduction-epilogue accumulate: tests, whole-model parity in all exact modes, smoke, bench lines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_bf16x3.py tests/test_model.py tests/test_gemm.py -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_call19_tests.log 2>&1; echo "tests rc=$?"; grep -n "^\[bf16x3\]\|^\[bf16x6\]\|passed\|failed\|^FAILED" gpurun_out/r2_call19_tests.log | cut -c1-200 | tail -70
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_call19_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r2_call19_smoke.log | cut -c1-300
for m in bf16x6 bf16x3 bf16; do
timeout 600 python bench.py --gemm $m --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c19_$m.json 2> gpurun_out/r2_bench_c19_$m.err; echo "bench $m rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_c19_$m.json') if l.startswith('{')][-1]); r=d['roofline']; print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['dtype'][:8], r.get('kernel'), r.get('kernel_ms_per_step'), d['config']['cuda_graph_error'])
" || tail -20 gpurun_out/r2_bench_c19_$m.err
done
