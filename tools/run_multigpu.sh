#!/bin/bash
# Multi-GPU round check (gpurun --gpus N -- 'bash tools/run_multigpu.sh N'): sharded image ==
# single-GPU image, then bench.py weak and strong scaling lines into gpurun_out/.
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 tools/check_multigpu.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP" | tail -4
$TR --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_n${N}_weak.json
$TR --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --scaling strong 2>/dev/null | tail -1 > gpurun_out/bench_n${N}_strong.json
for k in weak strong; do python -c "import json; d=json.load(open('gpurun_out/bench_n${N}_$k.json')); print('N=$N $k', d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['e2e']['value'], d['clocks'])"; done
