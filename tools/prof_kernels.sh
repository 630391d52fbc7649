#!/bin/bash
# ncu --set full of the two ASE kernels on ASE_medium-synth (one launch each) -> gpurun_out/<tag>_full.ncu-rep
TAG=${1:-prof}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"march_flat|integrate_ase_owner" -s 2 -c 2 \
    -o gpurun_out/${TAG}_full -f python tools/time_cases.py ASE_medium-synth > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_full.log
