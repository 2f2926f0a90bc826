#!/bin/bash
# N processes, one per GPU, running the host-link probe at the same time: is the link shared between GPUs?
#   tools/pcie_pair.sh [N=2] [pieces list, default "1 64"]
cd "$(dirname "$0")/.."
N=${1:-2}
for ((g=0; g<N; g++)); do CUDA_VISIBLE_DEVICES=$g python tools/pcie_probe2.py 1 > gpurun_out/pcie_pair_$g.txt 2>&1 & done
wait
for ((g=0; g<N; g++)); do echo "== GPU $g of $N (all running)"; cat gpurun_out/pcie_pair_$g.txt; done
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; cat gpurun_out/topo.txt; nproc; lscpu | grep -i "numa\|model name\|socket"
