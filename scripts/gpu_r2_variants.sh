#!/bin/bash
# A/B: the shipped library against variants built by benchmarks/build_variant.py (PCGMIX_LIB=...), RESIDENT timings
mkdir -p gpurun_out
V="pcgmix-a-data-augmentation-method-for-heart-sound-classification-extended_b200/csrc/variants"
for lib in "" $(ls $V/*.so 2>/dev/null); do
  echo "=== ${lib:-shipped}"
  for extra in "" "--stages 3" "--stages 5"; do
  PCGMIX_LIB=$lib timeout -s KILL 200 python benchmarks/run_configs.py --only resident --reps 100 $extra 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    if 'fused' in d['config']: print('   $extra', d['config'][:40], 'ms', round(d.get('ms_mean',0),4), 'frac', round(d.get('frac_of_measured_peak',0),4))
"
  done
done
