mkdir -p gpurun_out
N=${1:-2}
nvidia-smi topo -m > gpurun_out/r2_topo_$N.txt 2>&1
nproc >> gpurun_out/r2_topo_$N.txt; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> gpurun_out/r2_topo_$N.txt
cat /sys/fs/cgroup/cpu.max >> gpurun_out/r2_topo_$N.txt 2>&1
free -g >> gpurun_out/r2_topo_$N.txt
python benchmarks/probe_pcie_multi.py --out gpurun_out/r2_pcie_1of$N.json > /dev/null 2> gpurun_out/r2_pcie_1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 benchmarks/probe_pcie_multi.py --out gpurun_out/r2_pcie_$N.json > /dev/null 2> gpurun_out/r2_pcie_$N.err
python benchmarks/e2e_diag.py --out gpurun_out/r2_e2ediag_1of$N.json > /dev/null 2> gpurun_out/r2_e2ediag_1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 benchmarks/e2e_diag.py --out gpurun_out/r2_e2ediag_$N.json > /dev/null 2> gpurun_out/r2_e2ediag_$N.err
cat gpurun_out/r2_pcie_1of$N.json gpurun_out/r2_pcie_$N.json gpurun_out/r2_e2ediag_1of$N.json gpurun_out/r2_e2ediag_$N.json
tail -n 3 gpurun_out/r2_pcie_$N.err gpurun_out/r2_e2ediag_$N.err
