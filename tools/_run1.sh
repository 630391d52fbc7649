python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['kernel_ms_per_step'], d['roofline']['frac'], d['e2e']['image_time_ms'], d['gpu_launches'])"
