mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_tests.log
cat gpurun_out/r2_tests.log
timeout 120 python benchmarks/profile_augment_call.py > gpurun_out/r2_profile_call.txt 2>&1
head -3 gpurun_out/r2_profile_call.txt
timeout 120 python benchmarks/profile_augment_call.py --host-labels > gpurun_out/r2_profile_call_hostlabels.txt 2>&1
head -40 gpurun_out/r2_profile_call_hostlabels.txt
python examples/train_ddp_pcgmix.py --steps 60 > gpurun_out/r2_cfg5_1gpu.json 2> gpurun_out/r2_cfg5_1gpu.err
python examples/train_ddp_pcgmix.py --steps 60 --resident >> gpurun_out/r2_cfg5_1gpu.json 2>> gpurun_out/r2_cfg5_1gpu.err
cat gpurun_out/r2_cfg5_1gpu.json
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_a.json')); print(d['value'], d['roofline']['frac'], d['roofline']['serialized_launches']['frac'], d['e2e'])"; tail -3 gpurun_out/r2_bench_a.err
