# round 2: launch list of the bench command + full captures of the headline kernel (PCGmix+, float32 factor),
# its RESIDENT variant and the FLAT direct-load kernel of config 3 (each only after the plain command exited 0)
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-configs --no-variants --no-cfg5 --e2e-steps 3"
$B > gpurun_out/r2n_plain.json 2> gpurun_out/r2n_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_steps20.csv $B > gpurun_out/r2n_ncu0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mix_pipeline -s 10 -c 2 -o gpurun_out/r2_prof_headline -f $B --no-graph > gpurun_out/r2n_ncu1.log 2>&1
if [ "$1" = "all" ]; then
R="python benchmarks/run_configs.py --reps 20"
$R > gpurun_out/r2n_configs.jsonl 2> gpurun_out/r2n_configs.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:mix_kernel -s 12 -c 1 -o gpurun_out/r2_prof_cfg3_flat -f $R > gpurun_out/r2n_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mix_pipeline --launch-skip 12 -c 1 -o gpurun_out/r2_prof_resident -f python benchmarks/run_configs.py --only resident --reps 10 > gpurun_out/r2n_ncu3.log 2>&1
fi
tail -n 2 gpurun_out/r2n_ncu1.log
ls -la gpurun_out/*.ncu-rep
