for mb in 4 5; do RTB200_LIB=$PWD/raytrace-miniapp_b200/librtb200_mm$mb.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('march minblocks=$mb', d['ms_per_step'], d['kernel_ms_per_step'])"; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
