mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
nproc >> gpurun_out/topo.txt; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> gpurun_out/topo.txt
CUDA_VISIBLE_DEVICES=0 python benchmarks/probe_bandwidth.py > gpurun_out/bw0.json 2>/dev/null &
CUDA_VISIBLE_DEVICES=1 python benchmarks/probe_bandwidth.py > gpurun_out/bw1.json 2>/dev/null &
wait
cat gpurun_out/topo.txt; cat gpurun_out/bw0.json gpurun_out/bw1.json
