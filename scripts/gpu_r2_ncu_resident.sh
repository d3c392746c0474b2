#!/bin/bash
# ncu --set full of the RESIDENT PCGmix+ pipelined kernel (float32 spline)
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mix_pipeline --launch-skip 12 -c 1 \
    -o gpurun_out/r2_resident_v3_float32 -f python benchmarks/run_configs.py --only resident --reps 10 --spline float32 > gpurun_out/ncu_res3.log 2>&1
tail -2 gpurun_out/ncu_res3.log
