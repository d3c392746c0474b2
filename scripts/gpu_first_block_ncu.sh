#!/bin/bash
# ncu --set full capture of the first-block kernels (second launch of each: warm), after the plain run has exited 0
mkdir -p gpurun_out
timeout -s KILL 120 python benchmarks/first_block_profile.py > gpurun_out/fb_prof_plain.log 2>&1 || { tail -5 gpurun_out/fb_prof_plain.log; exit 1; }
timeout -s KILL 200 ncu --set full --clock-control none --import-source on -k regex:'apply_kernel|patch_moments_kernel|fold_kernel' \
    --launch-skip 3 --launch-count 5 -f -o gpurun_out/r2_prof_first_block python benchmarks/first_block_profile.py > gpurun_out/fb_ncu.log 2>&1
tail -3 gpurun_out/fb_ncu.log
ls -la gpurun_out/r2_prof_first_block.ncu-rep
