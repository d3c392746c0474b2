mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/h_plain.json 2> gpurun_out/h_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_headline.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/h_ncu0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mix_pipeline -s 10 -c 2 -o gpurun_out/prof_headline -f python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 3 --no-graph > gpurun_out/h_ncu1.log 2>&1
tail -n 2 gpurun_out/h_ncu1.log
