mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2_tests.log
cat gpurun_out/r2_tests.log
for prec in float32 float64; do
PCGMIX_SPLINE=$prec timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r2_bench_$prec.json 2> gpurun_out/r2_bench_$prec.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_$prec.json')); r=d['roofline']; print('$prec', 'value', round(d['value']/1e6,2), 'overlapped ms', round(r['kernel_ms_mean'],4), 'frac', round(r['frac'],3), 'serial ms', round(r['serialized_launches']['kernel_ms_mean'],4), 'min', round(r['serialized_launches']['kernel_ms_min'],4), round(r['serialized_launches']['frac'],3))"; tail -2 gpurun_out/r2_bench_$prec.err
done
timeout 200 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --e2e-steps 3 --method durratiomixup 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('pcgmix', 'value', round(d['value']/1e6,2), 'overlapped ms', round(r['kernel_ms_mean'],4), 'serial ms', round(r['serialized_launches']['kernel_ms_mean'],4), 'min', round(r['serialized_launches']['kernel_ms_min'],4))"
timeout 200 python benchmarks/run_configs.py --only resident --reps 100 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'][:60], round(d['ms_mean'],4), round(d.get('frac_of_measured_peak',0),3))"
