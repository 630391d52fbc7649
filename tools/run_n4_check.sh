#!/bin/bash
# 4-GPU check: the headline line (ASE_medium-synth, weak + strong + parity) and config 4 (S4 family).
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/n$N.err | tail -1 > gpurun_out/r02g_bench_n$N.json; echo "bench rc=$?"
timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --workload s4 --no-cpu-baseline 2> gpurun_out/n${N}_s4.err | tail -1 > gpurun_out/r02g_bench_s4_n$N.json; echo "bench s4 rc=$?"
timeout 200 python bench.py --steps 10 --warmup 3 --workload s4 --no-cpu-baseline 2> gpurun_out/n1_s4.err | tail -1 > gpurun_out/r02g_bench_s4_n1.json; echo "bench s4 n1 rc=$?"
for f in gpurun_out/r02g_bench_n$N.json gpurun_out/r02g_bench_s4_n$N.json gpurun_out/r02g_bench_s4_n1.json; do python -c "
import json,sys; d=json.load(open('$f'))
print('$f', d['value'], d['ms_per_step'], d['e2e']['image_time_ms'], d['roofline']['frac'], d.get('strong'), d.get('parity'))"; done
