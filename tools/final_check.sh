#!/bin/bash
# Final round check on the GPU box: the -m gpu suite and smoke(); when both pass, the profile round
# (bench line, ncu launch list, ncu --set full captures, every configuration) with tag $1.
TAG=${1:-r02i}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; T=$?
tail -n 15 gpurun_out/${TAG}_tests.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; S=$?
tail -n 2 gpurun_out/${TAG}_smoke.log
echo "tests rc=$T smoke rc=$S"
if [ $T -eq 0 ] && [ $S -eq 0 ]; then bash tools/profile_round.sh $TAG; fi
