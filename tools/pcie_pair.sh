#!/bin/bash
# two processes, one per GPU, running the host-link probe at the same time: is the link shared between GPUs?
cd "$(dirname "$0")/.."
for g in 0 1; do CUDA_VISIBLE_DEVICES=$g python tools/pcie_probe2.py > gpurun_out/pcie_pair_$g.txt 2>&1 & done
wait
for g in 0 1; do echo "== GPU $g (both running)"; cat gpurun_out/pcie_pair_$g.txt; done
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; cat gpurun_out/topo.txt
