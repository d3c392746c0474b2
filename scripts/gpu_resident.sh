mkdir -p gpurun_out
python -m pytest tests/test_resident_gpu.py tests/test_segment_gpu.py tests/test_segmentation_golden.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/res_tests.log
cat gpurun_out/res_tests.log
python benchmarks/run_configs.py --only resident > gpurun_out/res_bench.jsonl 2> gpurun_out/res_bench.err
cat gpurun_out/res_bench.jsonl; tail -5 gpurun_out/res_bench.err
