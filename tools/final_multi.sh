#!/bin/bash
# Round-end multi-GPU bench lines (torchrun, one rank per GPU) on the GPUs of this box: N = 8, 4, 2.
cd "$(dirname "$0")/.."
for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29530 + n)) bench.py --gpus $n --no-cpu-baseline > gpurun_out/r2_bench_n$n.json 2>gpurun_out/r2_bench_n$n.err
  python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n$n.json')); print($n, {k:d.get(k) for k in ('value','ms_per_step','eager_ms_per_step','graph_ms_per_step','fwd_ms','bwd_ms','host_enqueue_ms_per_step')}, d['e2e']['ms_per_step'])"
done
