#!/bin/bash
# A/B tuning: build librtb200 with extra nvcc flags into .variants/<name>.so (git-ignored, but it
# travels to the GPU box); select it with RTB200_LIB=.variants/<name>.so.
#   tools/build_variant.sh mb3 -DRTB_MARCH_MINBLOCKS=3
set -e
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$ROOT/.variants"
cd "$ROOT/raytrace-miniapp_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
     -Xcompiler -fPIC,-ffp-contract=off,-Wall -shared -cudart static "$@" \
     -o "$ROOT/.variants/$NAME.so" rtb200_kernels.cu rtb200_host.cu rtb200_multi.cu rtb200_dat.cpp -ldl
echo "$ROOT/.variants/$NAME.so"
