#!/bin/bash
cd "$(dirname "$0")/.."
N=${1:-2}
for thr in "" 2 0; do
  for ((g=0; g<N; g++)); do B200CTC_HOST_THREADS=$thr CUDA_VISIBLE_DEVICES=$g python tools/e2e_probe.py > gpurun_out/e2e_pair_$g.txt 2>&1 & done
  wait
  for ((g=0; g<N; g++)); do echo "== GPU $g of $N, B200CTC_HOST_THREADS='$thr'"; cat gpurun_out/e2e_pair_$g.txt; done
done
echo "== alone"; CUDA_VISIBLE_DEVICES=0 python tools/e2e_probe.py
