#!/bin/bash
# ncu --set full of the seeded path (march + scatter integration) on seed_small -> gpurun_out/<tag>_seed.ncu-rep
TAG=${1:-prof}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"integrate_seeded|integrate_scatter" -s 2 -c 1 \
    -o gpurun_out/${TAG}_seed -f python tools/time_cases.py seed_small > gpurun_out/${TAG}_ncu_seed.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_seed.log
