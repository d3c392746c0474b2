# 8-GPU box: sustained host<->device ceiling with 8 ranks, e2e loop pieces, cfg5 at 4 and 8 ranks
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_8.txt 2>&1
nproc >> gpurun_out/r2_topo_8.txt; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> gpurun_out/r2_topo_8.txt
free -g >> gpurun_out/r2_topo_8.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 4 8; do
timeout 120 $TR --nproc-per-node $n --master-port 2955$n benchmarks/probe_pcie_multi.py --seconds 0.5 --out gpurun_out/r2_pcie_$n.json > /dev/null 2> gpurun_out/r2_pcie_$n.err
done
timeout 150 $TR --nproc-per-node 8 --master-port 29561 benchmarks/e2e_diag.py --steps 12 --out gpurun_out/r2_e2ediag_8.json > /dev/null 2> gpurun_out/r2_e2ediag_8.err
for n in 4 8; do
timeout 120 $TR --nproc-per-node $n --master-port 2957$n examples/train_ddp_pcgmix.py --steps 60 > gpurun_out/r2_cfg5_${n}gpu.json 2> gpurun_out/r2_cfg5_${n}gpu.err
timeout 120 $TR --nproc-per-node $n --master-port 2958$n examples/train_ddp_pcgmix.py --steps 60 --resident >> gpurun_out/r2_cfg5_${n}gpu.json 2>> gpurun_out/r2_cfg5_${n}gpu.err
done
cat gpurun_out/r2_pcie_4.json gpurun_out/r2_pcie_8.json gpurun_out/r2_e2ediag_8.json gpurun_out/r2_cfg5_4gpu.json gpurun_out/r2_cfg5_8gpu.json
