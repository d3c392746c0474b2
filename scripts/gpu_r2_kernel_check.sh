# correctness of the pipelined kernel (mix + resident suites), CTA timeline and timing of both methods / precisions
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_mix_gpu.py tests/test_resident_gpu.py -m gpu -q -x 2>&1 | tail -3
for m in "durmixmagwarp(0.2,4)" durratiomixup; do
  PCGMIX_PROFILING_LIB=1 timeout 120 python benchmarks/cta_timeline.py --method "$m" | tail -1
  for prec in float32 float64; do
  [ "$m" = durratiomixup ] && [ $prec = float64 ] && continue
  timeout 120 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --e2e-steps 3 --no-configs --no-variants --no-cfg5 --method "$m" --spline $prec 2>gpurun_out/kc.err | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$m', '$prec', 'overlapped', round(r['kernel_ms_mean'],4), 'serial', round(r['serialized_launches']['kernel_ms_mean'],4), 'min', round(r['serialized_launches']['kernel_ms_min'],4), d['verified']['ok'])"
  done
done
tail -3 gpurun_out/kc.err
timeout 120 python benchmarks/run_configs.py --only resident --reps 100 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        if 'fused' in d['config']: print(d['config'][:40], round(d['ms_mean'],4), round(d.get('frac_of_measured_peak',0),3))"
timeout 120 python benchmarks/run_configs.py --only resident --reps 100 --spline float64 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)
        if 'fused' in d['config']: print('float64', d['config'][:40], round(d['ms_mean'],4), round(d.get('frac_of_measured_peak',0),3))"
