#!/bin/bash
mkdir -p gpurun_out
TFB_CAPTURE_DEBUG=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c9_bench.log 2>&1; grep -v Warning gpurun_out/r2_c9_bench.log | grep -B30 "Error" | head -60
TFB_CAPTURE_DEBUG=1 TFB_RAW_INPUTS=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c9_bench_noraw.log 2>&1; tail -c 600 gpurun_out/r2_c9_bench_noraw.log
