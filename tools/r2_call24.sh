#!/bin/bash
# call 24 (final tree, 1 GPU, last GPU seconds of the round): ncu launch list of ONE WHOLE step (cudaProfilerStart/Stop bracket: forward, backward, optimizer)
mkdir -p gpurun_out
timeout 170 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 1 --warmup 3 --graph 0 --no-cpu-baseline > gpurun_out/r2_ncu_bench_final.log 2>&1; echo "ncu step rc=$?"
python tools/ncu_step_summary.py gpurun_out/r2_launches_final.csv gpurun_out/r2_ncu_launch_summary_final.txt gpurun_out/r2_ncu_traffic_final.json | head -24
gzip -f gpurun_out/r2_launches_final.csv
