mkdir -p gpurun_out
python benchmarks/probe_bandwidth.py > gpurun_out/bw_probe.json 2>gpurun_out/bw_probe.err; cat gpurun_out/bw_probe.json
ncu --set full --clock-control none --import-source on -k regex:mix_resident -s 12 -c 1 -o gpurun_out/prof_resident_warp -f python benchmarks/run_configs.py --only resident --reps 20 > gpurun_out/ncu_res1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mix_resident -s 45 -c 1 -o gpurun_out/prof_resident_mix -f python benchmarks/run_configs.py --only resident --reps 20 > gpurun_out/ncu_res2.log 2>&1
tail -3 gpurun_out/ncu_res1.log gpurun_out/ncu_res2.log
