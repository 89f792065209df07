#!/bin/bash
# Round-2 first GPU call: state of the whole -m gpu suite (no -x, xfails run), sidecar noise, per-kernel launch list.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --runxfail -p no:cacheprovider > gpurun_out/r2_call1_gpu_tests.log 2>&1
tail -40 gpurun_out/r2_call1_gpu_tests.log
timeout 300 python tools/sidecar_bisect.py > gpurun_out/r2_sidecar_bisect.log 2>&1
cat gpurun_out/r2_sidecar_bisect.log | tail -8
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench0.json 2> gpurun_out/r2_bench0.err
cut -c1-400 gpurun_out/r2_bench0.json
timeout 240 python tools/gemm_bench.py > gpurun_out/r2_gemm_bench.log 2>&1
tail -30 gpurun_out/r2_gemm_bench.log
# launch list of one eager step (per-kernel GPU time, serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches0.csv python bench.py --steps 1 --warmup 3 --graph 0 --no-cpu-baseline > gpurun_out/r2_ncu_bench.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(l for l in open('gpurun_out/r2_launches0.csv') if l.startswith('"'))]
h = rows[0]; ki, vi = h.index('Kernel Name'), h.index('Metric Value'); ui = h.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(',', '')); v = v / 1e3 if r[ui] in ('ns', 'nsecond') else v
    agg[r[ki][:70]][0] += 1; agg[r[ki][:70]][1] += v
tot = sum(v[1] for v in agg.values())
print('total us', tot, 'launches', sum(v[0] for v in agg.values()))
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print('%9.1f us %5.1f%% x%-5d %s' % (us, 100 * us / tot, n, k))
PY
