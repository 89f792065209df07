#!/bin/bash
# Round-2 GPU call 5: SE backward rewrite, decoder streams, wgrad CTA caps; then the ncu evidence passes.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_ops.py tests/test_trainer.py -m gpu -q -x -k "se or trainer or graph" -p no:cacheprovider > gpurun_out/r2_c5_new_tests.log 2>&1; echo "new tests rc=$?"; tail -3 gpurun_out/r2_c5_new_tests.log | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_c5_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_c5_smoke.log | cut -c1-300
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c5_$name.json 2> gpurun_out/r2_bench_c5_$name.err; }
run default TFB_X=0
run decoder_one_stream TFB_DECODER_STREAMS=0
run wgrad_cap32 TFB_WGRAD_MAX_CTAS=32
run wgrad_cap64 TFB_WGRAD_MAX_CTAS=64
run wgrad_cap96 TFB_WGRAD_MAX_CTAS=96
run single_stream TFB_TWO_STREAMS=0 TFB_WGRAD_STREAM=0
for f in gpurun_out/r2_bench_c5_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); r=d['roofline']; print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], r.get('kernel_ms_per_step'), r.get('frac'), r.get('roof_frac'), r.get('large_launches',{}).get('achieved'), d['config']['cuda_graph_error'])
except Exception as e: print('ERR', e); print(open('$f'.replace('.json','.err')).read()[-1200:])
"; done
# ---- ncu evidence: (1) every hot kernel once with the full section set, (2) DRAM traffic of the GEMM launches of one step, (3) launch list of one step
WARM=0 timeout 1200 ncu --set full --clock-control none --import-source on -f -o gpurun_out/r2_kernels python tools/ncu_targets.py > gpurun_out/r2_ncu_targets.log 2>&1; echo "ncu targets rc=$?"; tail -2 gpurun_out/r2_ncu_targets.log
ls -la gpurun_out/r2_kernels.ncu-rep
timeout 600 ncu -i gpurun_out/r2_kernels.ncu-rep --page raw --csv > gpurun_out/r2_kernels_raw.csv 2> /dev/null
python tools/ncu_summarize.py gpurun_out/r2_kernels_raw.csv > gpurun_out/r2_ncu_kernels.txt 2> gpurun_out/r2_ncu_summarize.err; head -5 gpurun_out/r2_ncu_kernels.txt; wc -l gpurun_out/r2_ncu_kernels.txt
gzip -f gpurun_out/r2_kernels_raw.csv
if [ $(stat -c %s gpurun_out/r2_kernels.ncu-rep) -gt 45000000 ]; then rm gpurun_out/r2_kernels.ncu-rep; echo "ncu-rep too large for the pull limit: kept the raw csv only"; fi
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 5100 -c 1750 --csv --log-file gpurun_out/r2_launches_step.csv python bench.py --steps 1 --warmup 3 --graph 0 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1; echo "ncu step rc=$?"
python - <<'PY'
import csv, collections, json
rows = [r for r in csv.reader(l for l in open('gpurun_out/r2_launches_step.csv') if l.startswith('"'))]
h = rows[0]; ki, mi, vi, ui = h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(',', '')); k = r[ki].replace('void ', '').replace('<unnamed>::', '').split('(')[0][:60]
    if r[mi] == 'gpu__time_duration.sum':
        agg[k][0] += 1; agg[k][1] += v / 1e3 if r[ui] in ('ns', 'nsecond') else v
    else:
        agg[k][2] += v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(r[ui], 1)
tot = sum(v[1] for v in agg.values())
with open('gpurun_out/r2_ncu_launch_summary.txt', 'w') as f:
    f.write('one eager training step (batch 10, bf16 mode) under ncu --metrics gpu__time_duration.sum,dram__bytes_* (serialised, cold caches): %d launches, %.2f ms of kernel time\n' % (sum(v[0] for v in agg.values()), tot / 1e3))
    for k, (n, us, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write('%9.1f us %5.1f%% x%-5d %9.2f MB DRAM/launch  %s\n' % (us, 100 * us / tot, n, by / max(n, 1) / 1e6, k))
print(open('gpurun_out/r2_ncu_launch_summary.txt').read()[:3500])
g = [(k, v) for k, v in agg.items() if 'gemm_tc_kernel' in k]
n = sum(v[0] for _, v in g); by = sum(v[2] for _, v in g)
json.dump({'gemm_tc_kernel': {'dram_bytes_per_launch': round(by / max(n, 1)), 'launches': n, 'source': 'ncu dram__bytes_read.sum + dram__bytes_write.sum over the %d gemm_tc_kernel launches of one eager step (profiles/r2_ncu_launch_summary.txt)' % n}}, open('gpurun_out/r2_ncu_traffic.json', 'w'))
PY
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_call5_gpu_tests.log 2>&1; echo "suite rc=$?"; tail -6 gpurun_out/r2_call5_gpu_tests.log | cut -c1-300
