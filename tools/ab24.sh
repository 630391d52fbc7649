mkdir -p gpurun_out
timeout 300 python tools/ab_variants.py tree segbase tree segbase > gpurun_out/ab24.log 2>&1; cat gpurun_out/ab24.log
RTB200_LIB=$PWD/.variants/segbase.so timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_configs.py -x -q 2>&1 | tail -n 3
