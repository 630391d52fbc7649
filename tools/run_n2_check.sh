mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/n2_tests.log 2>&1; echo "multi tests rc=$?"; tail -n 4 gpurun_out/n2_tests.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/n2.err | tail -1 > gpurun_out/r02g_bench_n2.json; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02g_bench_n2.json'))
print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['e2e'], d['strong'], d['parity'])"
timeout 200 python tools/time_multi.py > gpurun_out/n2_multi.log 2>&1; cat gpurun_out/n2_multi.log
