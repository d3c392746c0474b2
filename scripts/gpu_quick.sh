mkdir -p gpurun_out
python -m pytest tests/test_mix_gpu.py -m gpu -x -q 2>&1 | tail -2
for opt in "" "--batch 16384 --resident-batches 2"; do
python bench.py --steps 60 --warmup 5 --no-cpu-baseline --e2e-steps 3 $opt 2>gpurun_out/q.err | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$opt'.ljust(48), 'overlapped ms', round(r['kernel_ms_mean'],4), 'frac', round(r['frac'],3), 'serial ms', round(r['serialized_launches']['kernel_ms_mean'],4), round(r['serialized_launches']['frac'],3), 'value', round(d['value']/1e6,1))"
done
timeout 100 python benchmarks/run_configs.py --only resident --reps 100 --stages 4 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'][:40], round(d['ms_mean'],4), round(d.get('frac_of_measured_peak',0),3))"
tail -n 3 gpurun_out/q.err
