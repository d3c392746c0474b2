#!/bin/bash
# first conv block: GPU tests, then its bench legs (one GPU)
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_first_block.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/fb_tests.log
cat gpurun_out/fb_tests.log
timeout -s KILL 120 python benchmarks/run_configs.py --only first_block > gpurun_out/fb_bench.jsonl 2> gpurun_out/fb_bench.err
cat gpurun_out/fb_bench.jsonl | cut -c1-400
tail -3 gpurun_out/fb_bench.err
