set -x; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/a_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1
python bench.py > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/a_ref.json 2> gpurun_out/a_ref.err
python benchmarks/run_configs.py > gpurun_out/a_configs.jsonl 2> gpurun_out/a_configs.err
cat gpurun_out/a_tests.log gpurun_out/a_smoke.log
