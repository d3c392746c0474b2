mkdir -p gpurun_out
: > gpurun_out/res_sweep.txt
timeout 300 python -m pytest tests/test_resident_gpu.py -m gpu -x -q 2>&1 | tail -3 >> gpurun_out/res_sweep.txt
for opt in "" "--stages 4" "--stages 2"; do
  echo "== $opt" >> gpurun_out/res_sweep.txt
  timeout 120 python benchmarks/run_configs.py --only resident --reps 100 $opt 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['config'][:40], round(d['ms_mean'],4), round(d.get('frac_of_measured_peak',0),3))
    elif 'rror' in l: print(l.strip()[:200])
" >> gpurun_out/res_sweep.txt
done
cat gpurun_out/res_sweep.txt
