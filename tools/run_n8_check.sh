#!/bin/bash
# 8-GPU round check: the weak-scaling bench line (which carries the strong-scaling and parity
# sub-records) and the C ABI's multi-device entry (tools/time_multi.py).
N=${1:-8}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/n$N.err | tail -1 > gpurun_out/r02g_bench_n$N.json; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02g_bench_n$N.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['strong'], d['parity'], d['roofline']['frac'])"
timeout 200 python tools/time_multi.py > gpurun_out/n${N}_multi.log 2>&1; cat gpurun_out/n${N}_multi.log
