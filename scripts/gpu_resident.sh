mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_resident_gpu.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/res_tests.log
cat gpurun_out/res_tests.log
timeout 200 python benchmarks/run_configs.py --only resident > gpurun_out/res_bench.jsonl 2> gpurun_out/res_bench.err
cat gpurun_out/res_bench.jsonl; tail -5 gpurun_out/res_bench.err
