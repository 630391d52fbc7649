# A/B of the march kernel's occupancy variants (.variants/*.so built by tools/build_variant.sh)
python -m pytest tests/test_gpu_parity.py tests/test_gpu_medium.py tests/test_gpu_seeded.py -x -q -m gpu 2>&1 | tail -5
echo default; python tools/time_cases.py 2>&1 | tail -3
for v in "$@"; do echo $v; RTB200_LIB=.variants/$v.so python tools/time_cases.py 2>&1 | tail -3; done
