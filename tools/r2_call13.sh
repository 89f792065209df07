#!/bin/bash
# 8-GPU weak-scaling run of the headline bench (charged 8x: keep it to one bench invocation)
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "rc=$?"
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_n8.json') if l.startswith('{')][-1]); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['cuda_graph'], d['config']['cuda_graph_error'])"
