#!/bin/bash
# call 20 (2 GPUs): readiness-aware exchange spans (GradAllReducer.replan), span count sweep at N=2; N=1 before / after
mkdir -p gpurun_out
run2() { name=$1; port=$2; shift; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_c20_$name.json 2> gpurun_out/r2_bench_c20_$name.err; }
run1() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c20_$name.json 2> gpurun_out/r2_bench_c20_$name.err; }
run2 n2_default 29541 TFB_X=0
run2 n2_chunks4 29542 TFB_GRAD_CHUNKS=4
run2 n2_chunks12 29543 TFB_GRAD_CHUNKS=12
run2 n2_nosplit8 29544 TFB_SPLIT_LATE_SPANS=0
run1 n1_default TFB_X=0
run1 n1_old TFB_SPLIT_LATE_SPANS=0 TFB_GRAD_CHUNKS=24
for f in gpurun_out/r2_bench_c20_*.json; do echo $f; python -c "
import json
try:
    d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['config']['cuda_graph'], d['config']['cuda_graph_error'])
except Exception as e: print('ERR', e); print(open('$f'.replace('.json','.err')).read()[-800:])
"; done
