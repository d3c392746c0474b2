mkdir -p gpurun_out
python -m pytest tests/test_mix_gpu.py -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 3 > gpurun_out/q_warp.json 2>gpurun_out/q.err
python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 3 --method durratiomixup > gpurun_out/q_mix.json 2>>gpurun_out/q.err
python - <<'PY'
import json
for f in ('q_warp','q_mix'):
    d=json.load(open(f'gpurun_out/{f}.json')); r=d['roofline']
    print(f, 'overlapped ms', round(r['kernel_ms_mean'],4), 'frac', round(r['frac'],3), 'serial ms', round(r['serialized_launches']['kernel_ms_mean'],4), round(r['serialized_launches']['frac'],3), 'value', round(d['value']/1e6,1))
PY
tail -3 gpurun_out/q.err
