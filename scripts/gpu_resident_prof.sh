mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_resident.csv python benchmarks/run_configs.py --only resident --reps 20 > gpurun_out/ncu_res0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mix_pipeline -s 12 -c 1 -o gpurun_out/prof_respipe_warp -f python benchmarks/run_configs.py --only resident --reps 20 > gpurun_out/ncu_res1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mix_pipeline -s 95 -c 1 -o gpurun_out/prof_respipe_mix -f python benchmarks/run_configs.py --only resident --reps 20 > gpurun_out/ncu_res2.log 2>&1
tail -n 3 gpurun_out/ncu_res1.log gpurun_out/ncu_res2.log
