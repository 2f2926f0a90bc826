#!/bin/bash
# Round-end evidence on one GPU: tests, smoke, bench line, reference arm, ncu launch list + full captures, sweep, LN comparison.
cd "$(dirname "$0")/.."
timeout -s KILL 900 python -m pytest tests -q -m gpu --timeout=300 -p no:cacheprovider > gpurun_out/pytest_final.log 2>&1; tail -3 gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_ref_n1.json 2>gpurun_out/r2_ref_n1.err
python bench.py > gpurun_out/r2_bench_n1.json 2>gpurun_out/r2_bench_n1.err; tail -2 gpurun_out/r2_bench_n1.err
python tools/sweep.py > gpurun_out/r2_sweep.txt 2>&1
python tools/ln_bench.py > gpurun_out/r2_ln_bench.txt 2>&1; tail -3 gpurun_out/r2_ln_bench.txt
A="--steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-graph --no-kernel-loop"
python bench.py $A > gpurun_out/plain_r2.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 20 -c 30 --csv --log-file gpurun_out/r2_launches.csv python bench.py $A > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"lattice_kernel|gradient_ring|softmax_gather_ring" -s 9 -c 3 -o gpurun_out/prof_r2 -f python bench.py $A > gpurun_out/ncu_f.log 2>&1; tail -1 gpurun_out/ncu_f.log
ncu --set full --clock-control none --import-source on -k regex:"ln_softmax_gather|ln_gradient" -s 2 -c 2 -o gpurun_out/prof_r2_ln -f python tools/ln_bench.py > gpurun_out/ncu_ln.log 2>&1; tail -1 gpurun_out/ncu_ln.log
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','eager_ms_per_step','fwd_ms','bwd_ms','host_enqueue_ms_per_step')}, d['roofline']['frac'], d['roofline']['kernel_ms'], d['step_roofline']['frac'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'], d['clocks'])"
