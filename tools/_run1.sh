python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; tail -2 gpurun_out/bench_r1_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1_n1.json')); print(d['value'], d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'], d['e2e'], d['cpu_baseline']['value'], d['cpu_baseline']['cores'], d['clocks'])"
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-260
python __graft_entry__.py smoke
python tools/time_cases.py
