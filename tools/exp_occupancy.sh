# A/B of march kernel variants (.variants/*.so built by tools/build_variant.sh)
echo default; python tools/time_cases.py 2>&1 | tail -3
for v in "$@"; do echo $v; RTB200_LIB=.variants/$v.so python tools/time_cases.py 2>&1 | tail -3; done
