#!/bin/bash
# two-GPU box: the tests that need two devices, then the 2-rank bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "two_devices or two_gpu or multi" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/a_bench_2gpu.json 2> gpurun_out/a_bench_2gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/a_bench_2gpu.json').read().strip().splitlines()[-1])
print(round(d['value'],1), d['ms_per_step'], d.get('verified',{}).get('ok'), round(d['e2e']['value']), d.get('cfg5',{}).get('padded_batches',{}).get('step_ms_max_over_ranks'))
PY
