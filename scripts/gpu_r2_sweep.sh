# usage: gpu_r2_sweep.sh "1 2" | "4 8"   — the 1 M-cycle sweep at the given rank counts
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in $1; do
  if [ "$n" = "1" ]; then
    timeout 300 python benchmarks/sweep_1m.py >> gpurun_out/r2_final_sweep_1m.jsonl 2> gpurun_out/r2_final_sweep_$n.err
  else
    timeout 300 $TR --nproc-per-node $n --master-port 2962$n benchmarks/sweep_1m.py >> gpurun_out/r2_final_sweep_1m.jsonl 2> gpurun_out/r2_final_sweep_$n.err
  fi
  tail -n 2 gpurun_out/r2_final_sweep_$n.err | cut -c1-200
done
grep cfg4 gpurun_out/r2_final_sweep_1m.jsonl
