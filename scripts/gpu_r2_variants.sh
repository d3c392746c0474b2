#!/bin/bash
# A/B: the shipped library against variants built by benchmarks/build_variant.py (PCGMIX_LIB=...), config 2 timings
mkdir -p gpurun_out
V="pcgmix-a-data-augmentation-method-for-heart-sound-classification-extended_b200/csrc/variants"
for lib in "" $(ls $V/*.so 2>/dev/null); do
  echo "=== ${lib:-shipped}"
  for prec in float32 float64; do
  PCGMIX_LIB=$lib timeout 120 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --e2e-steps 3 --no-configs --no-variants --no-cfg5 --spline $prec 2>gpurun_out/kc.err | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']; print('  pcgmix+ $prec', 'overlapped', round(r['kernel_ms_mean'],4), 'serial', round(r['serialized_launches']['kernel_ms_mean'],4), 'min', round(r['serialized_launches']['kernel_ms_min'],4), d['verified']['ok'])"
  done
done
