mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  ( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err
else
  ( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/r2_bench_full_$N.json 2> gpurun_out/r2_bench_full_$N.err
fi
tail -5 gpurun_out/r2_bench_full*.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_bench_full*.json')):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f,'unparsable',e); continue
    print(f, 'value', round(d['value']/1e6,2), 'frac', round(d['roofline']['frac'],3), 'serial', round(d['roofline']['serialized_launches']['frac'],3), 'verified', d['verified'])
    print(' e2e', round(d['e2e']['value']/1e6,3), d['e2e']['ms_per_step'], {k:(round(v['value']/1e6,3), round(v['ms_per_step'],3)) for k,v in d['e2e_variants'].items()})
    print(' overlapped', d['gpu_launches_overlapped'], 'cpu', d.get('cpu_baseline'))
    for c in d.get('configs',[]): print('   ', c['config'][:70], round(c['ms_mean'],4), round(c.get('frac_of_measured_peak',0),3))
    print(' cfg5', d.get('cfg5'))
PY
