#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/graph_vs_eager.py default "no dropout" DECODER TWO_STREAMS > gpurun_out/r2_graph_vs_eager.log 2>&1; echo "rc=$?"; cat gpurun_out/r2_graph_vs_eager.log | tail -15
