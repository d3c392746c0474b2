mkdir -p gpurun_out
: > gpurun_out/res_sweep.txt
for opt in "" "--stages 3" "--stages 5" "--consumer-threads 448" "--pbuf-pct 45 --stages 5" "--max-slice 1252 --stages 6" "--max-slice 1252 --stages 8"; do
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
