# pipelined march / integration: ranges x resident march CTAs per SM
python tools/time_cases.py ASE_medium-synth 2>&1 | tail -1
for r in 4 8 16; do for m in 1 2 3; do echo "ranges $r march_ctas $m"; RTB200_PIPE_RANGES=$r RTB200_PIPE_MARCH_CTAS=$m python tools/time_cases.py ASE_medium-synth 2>&1 | tail -1; done; done
RTB200_PIPE_RANGES=8 RTB200_PIPE_MARCH_CTAS=2 python -m pytest tests/test_gpu_parity.py tests/test_gpu_medium.py -x -q -m gpu 2>&1 | tail -2
