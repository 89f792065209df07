#!/bin/bash
# call 22 (final tree, 1 GPU): ncu launch list of one step, full -m gpu suite, smoke, headline bench (with cpu baseline), other BASELINE configs, parity-mode line
mkdir -p gpurun_out
timeout 900 ncu --nvtx --nvtx-include "tfb_profiled_step/" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 1 --warmup 3 --graph 0 --no-cpu-baseline > gpurun_out/r2_ncu_bench_final.log 2>&1; echo "ncu step rc=$?"
python tools/ncu_step_summary.py gpurun_out/r2_launches_final.csv gpurun_out/r2_ncu_launch_summary_final.txt gpurun_out/r2_ncu_traffic_final.json | head -30 && cp gpurun_out/r2_ncu_traffic_final.json profiles/r2_ncu_traffic.json
gzip -f gpurun_out/r2_launches_final.csv
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_gpu_tests_final.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/r2_gpu_tests_final.log | cut -c1-300
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_smoke_final.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_smoke_final.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
for c in 3 4 5; do timeout 400 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_final_config$c.json 2> gpurun_out/r2_bench_final_config$c.err; echo "config $c rc=$?"; done
timeout 400 python bench.py --gemm bf16x6 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_final_bf16x6.json 2> gpurun_out/r2_bench_final_bf16x6.err; echo "x6 rc=$?"
for f in gpurun_out/r2_bench_final*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); r=d['roofline']; print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], r.get('kernel'), r.get('kernel_ms_per_step'), r.get('frac'), r.get('roof_frac'), r.get('traffic'), (d.get('cpu_baseline') or {}).get('value'), d['clocks'], d['config']['cuda_graph_error'])
except Exception as e: print('ERR', e); print(open('$f'.replace('.json','.err')).read()[-1200:])
"; done
