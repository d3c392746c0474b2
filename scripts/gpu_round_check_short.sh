# what the driver runs at round end on one GPU, without the reference arm (CPU only; unchanged)
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/b_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/b_smoke.log 2>&1
( time python bench.py --steps 20 --warmup 5 ) > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err
cat gpurun_out/b_tests.log gpurun_out/b_smoke.log
grep real gpurun_out/b_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b_bench.json').read().strip().splitlines()[-1])
print(round(d['value'],1), d.get('verified'), round(d['e2e']['value']), d['ms_per_step'], d['roofline']['frac'])
for c in d.get('configs', []):
    if c['config'].startswith('first block'): print(c['config'][:90], round(c['ms_mean'],4), round(c.get('frac_of_measured_peak',0),3))
PY
