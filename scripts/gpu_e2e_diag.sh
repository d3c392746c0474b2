mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 24 > gpurun_out/d1.json 2> gpurun_out/d1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 3 --e2e-steps 24 > gpurun_out/d3.json 2> gpurun_out/d3.err
for f in d1 d3; do python - $f <<'PY'
import sys,json
f=sys.argv[1]
raw=open(f'gpurun_out/{f}.json').read()
print(f, 'stdout lines:', raw.count('\n'))
d=json.loads(raw); e=d['e2e']; print(f,'e2e ms/step', round(e['ms_per_step'],3), 'host ms', round(e['host_ms_per_step_inside_augment'],3), 'draws ms', round(e['host_draws_ms_per_step'],3))
PY
done
