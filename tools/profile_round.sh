#!/bin/bash
# Round profile on the GPU box (run under gpurun from the repo root):
#   1. bench.py without a profiler (the number that counts), 2. the ncu launch list of the same
#   command, 3. one `--set full` capture of the two ASE kernels.  Outputs go to gpurun_out/;
#   tools/ncu_summary.py / ncu_lines.py / ncu_regions.py turn the report into profiles/*.txt.
set -e
TAG=${1:-r01_v9}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 2>gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench_n1.json
cat gpurun_out/${TAG}_bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"march_flat|integrate_ase_owner" -s 2 -c 2 \
    -o gpurun_out/${TAG}_full -f python tools/time_cases.py ASE_medium-synth > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_full.log
