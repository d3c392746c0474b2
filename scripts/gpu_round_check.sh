# what the driver runs at round end, plus the 2-rank bench when the box has two GPUs
set -x; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/a_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/a_ref.json 2> gpurun_out/a_ref.err
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
if [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/a_bench_2gpu.json 2> gpurun_out/a_bench_2gpu.err
fi
cat gpurun_out/a_tests.log gpurun_out/a_smoke.log
grep real gpurun_out/a_bench.err
python - <<'PY'
import json,glob
for f in ('gpurun_out/a_ref.json','gpurun_out/a_bench.json','gpurun_out/a_bench_2gpu.json'):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,'unparsable',e); continue
    print(f, round(d['value'],1), d.get('cpu_baseline',{}).get('kind'), d.get('verified'), round(d['e2e']['value']), d['ms_per_step'])
PY
