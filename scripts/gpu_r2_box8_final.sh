# 8-GPU box, current library: bench.py at 8 and 4 ranks (with cfg5 and the e2e variants), the 1 M-cycle sweep at 1/2/4/8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4; do
  timeout 400 $TR --nproc-per-node $n --master-port 2961$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_final_bench_${n}gpu.json 2> gpurun_out/r2_final_bench_${n}gpu.err
done
: > gpurun_out/r2_final_sweep_1m.jsonl
timeout 300 python benchmarks/sweep_1m.py >> gpurun_out/r2_final_sweep_1m.jsonl 2> gpurun_out/r2_final_sweep.err
for n in 2 4 8; do
  timeout 300 $TR --nproc-per-node $n --master-port 2962$n benchmarks/sweep_1m.py >> gpurun_out/r2_final_sweep_1m.jsonl 2>> gpurun_out/r2_final_sweep.err
done
cat gpurun_out/r2_final_sweep_1m.jsonl
python - <<'PY'
import json
for n in (8,4):
    try:
        d=json.loads(open(f'gpurun_out/r2_final_bench_{n}gpu.json').read().strip().splitlines()[-1])
        print(n,'value',round(d['value']/1e6,1),'e2e',round(d['e2e']['value']/1e6,3), {k:round(v['value']/1e6,2) for k,v in d['e2e_variants'].items()}, 'verified', d['verified']['ok'])
        print('  cfg5', {k:(v['step_ms_max_over_ranks'], v['augment_call_host_ms_median']) for k,v in d['cfg5'].items()})
    except Exception as e: print(n,'failed',e)
PY
tail -3 gpurun_out/r2_final_bench_8gpu.err gpurun_out/r2_final_sweep.err
