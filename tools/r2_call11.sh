#!/bin/bash
# 2-GPU check of the data-parallel path (NCCL all-reduce from the launch stream, pipelined AdamW) — keep it short (charged 2x)
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "rc=$?"
grep -v Warning gpurun_out/r2_bench_n2.err | tail -5 | cut -c1-300
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_n2.json') if l.startswith('{')][-1]); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['cuda_graph'], d['config']['cuda_graph_error'])"
