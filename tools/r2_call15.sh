#!/bin/bash
# call 15: per-process critical-path ablation, per-shape in-graph kernel times
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops.py -m gpu -q -x -p no:cacheprovider > gpurun_out/r2_call15_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_call15_tests.log | cut -c1-300
timeout 400 python tools/shape_times.py --top 90 > gpurun_out/r2_shape_times.txt 2> gpurun_out/r2_shape_times.err; echo "shapes rc=$?"; head -50 gpurun_out/r2_shape_times.txt | cut -c1-200; tail -3 gpurun_out/r2_shape_times.err
timeout 900 python tools/ablate_step.py > gpurun_out/r2_ablation.txt 2> gpurun_out/r2_ablation.err; echo "ablate rc=$?"; cat gpurun_out/r2_ablation.txt; tail -5 gpurun_out/r2_ablation.err
