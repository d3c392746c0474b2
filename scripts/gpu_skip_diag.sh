mkdir -p gpurun_out
: > gpurun_out/skip.txt
for opt in "" "--debug-skip 2" "--debug-skip 8" "--debug-skip 16" "--debug-skip 24" "--debug-skip 10" "--method durratiomixup"; do
  python bench.py --steps 100 --warmup 5 --no-cpu-baseline --e2e-steps 3 $opt 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$opt'.ljust(24), 'overlapped', round(r['kernel_ms_mean'],4), 'serial', round(r['serialized_launches']['kernel_ms_mean'],4))" >> gpurun_out/skip.txt
done
cat gpurun_out/skip.txt
