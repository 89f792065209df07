#!/bin/bash
# call 18 (2 GPUs): where does the data-parallel overhead come from? N=2 step with span count / NCCL CTA budget / no exchange
mkdir -p gpurun_out
run() { name=$1; port=$2; shift; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_c18_$name.json 2> gpurun_out/r2_bench_c18_$name.err; }
run default 29531 TFB_X=0
run noexchange 29532 TFB_NO_ALLREDUCE=1
run chunks96 29533 TFB_GRAD_CHUNKS=96
run chunks8 29534 TFB_GRAD_CHUNKS=8
run maxctas8 29535 NCCL_MAX_CTAS=8
run chunks96_maxctas8 29536 TFB_GRAD_CHUNKS=96 NCCL_MAX_CTAS=8
for f in gpurun_out/r2_bench_c18_*.json; do echo $f; python -c "
import json
try:
    d=json.loads([l for l in open('$f') if l.startswith('{')][-1]); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['cuda_graph'], d['config']['cuda_graph_error'])
except Exception as e: print('ERR', e); print(open('$f'.replace('.json','.err')).read()[-800:])
"; done
