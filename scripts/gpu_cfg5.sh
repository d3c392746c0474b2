mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  python examples/train_ddp_pcgmix.py --steps 60 > gpurun_out/cfg5_1gpu.json 2> gpurun_out/cfg5_1gpu.err
  python examples/train_ddp_pcgmix.py --steps 60 --resident >> gpurun_out/cfg5_1gpu.json 2>> gpurun_out/cfg5_1gpu.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     examples/train_ddp_pcgmix.py --steps 60 > gpurun_out/cfg5_${N}gpu.json 2> gpurun_out/cfg5_${N}gpu.err
fi
tail -n 2 gpurun_out/cfg5_${N}gpu.json; tail -n 3 gpurun_out/cfg5_${N}gpu.err
