#!/bin/bash
# Round-2 GPU call 6: pipelined AdamW, raw-input pipeline in the step, trainer tests; A/B; light ncu pass for the kernels call 5 missed.
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_trainer.py tests/test_pipeline.py -m gpu -q -p no:cacheprovider -s > gpurun_out/r2_c6_new_tests.log 2>&1; echo "new tests rc=$?"; grep -E "eager vs eager|passed|failed|Error" gpurun_out/r2_c6_new_tests.log | cut -c1-400
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c6_$name.json 2> gpurun_out/r2_bench_c6_$name.err; }
run default TFB_X=0
run adamw_tail TFB_OVERLAP_ADAMW=0
run expanded_inputs TFB_RAW_INPUTS=0
run wgrad_cap120 TFB_WGRAD_MAX_CTAS=120
run chunks8 TFB_GRAD_CHUNKS=8
env TFB_X=0 timeout 300 python bench.py --config 4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c6_config4.json 2> gpurun_out/r2_bench_c6_config4.err
env TFB_X=0 timeout 300 python bench.py --config 5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_c6_config5.json 2> gpurun_out/r2_bench_c6_config5.err
for f in gpurun_out/r2_bench_c6_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.load(open('$f')); r=d['roofline']; print(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['h2d_bytes_per_step'], d['gpu_launches'], r.get('kernel_ms_per_step'), r.get('frac'), r.get('roof_frac'), r.get('traffic'), d['config']['cuda_graph_error'])
except Exception as e: print('ERR', e); print(open('$f'.replace('.json','.err')).read()[-1200:])
"; done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active,lts__t_sectors.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__throughput.avg.pct_of_peak_sustained_elapsed
WARM=0 SECTIONS=attn,decoder,ln,bev,adamw timeout 600 ncu --metrics $M --clock-control none --page raw --csv --log-file gpurun_out/r2_kernels_part2_raw.csv python tools/ncu_targets.py > gpurun_out/r2_ncu_targets2.log 2>&1; echo "ncu part2 rc=$?"
python tools/ncu_summarize.py gpurun_out/r2_kernels_part2_raw.csv > gpurun_out/r2_ncu_kernels_part2.txt 2> gpurun_out/r2_ncu_summarize2.err; grep -E "attn|adamw|bev_|upsample|conv3x3" gpurun_out/r2_ncu_kernels_part2.txt | cut -c1-170 | head -30
WARM=0 SECTIONS=gpt timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -c 4 -f -o gpurun_out/r2_full_gemm python tools/ncu_targets.py > gpurun_out/r2_ncu_full_gemm.log 2>&1; echo "ncu full gemm rc=$?"
WARM=0 SECTIONS=attn timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_tc_kernel -c 2 -f -o gpurun_out/r2_full_attn python tools/ncu_targets.py > gpurun_out/r2_ncu_full_attn.log 2>&1; echo "ncu full attn rc=$?"
ls -la gpurun_out/*.ncu-rep
