#!/bin/bash
# Round profile on the GPU box (run under gpurun from the repo root):
#   1. bench.py without a profiler (the number that counts), 2. the ncu launch list of the same
#   command (time-only pass), 3. `--set full` captures of the ASE kernels (ASE_medium-synth) and of
#   the seeded kernels (seed_small), 4. every BASELINE configuration (tools/run_configs.py).
# Outputs go to gpurun_out/; tools/ncu_summary.py / ncu_traffic.py / ncu_regions.py turn the
# reports into profiles/*.
TAG=${1:-r02}
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 2>gpurun_out/${TAG}_bench.err | tail -1 > gpurun_out/${TAG}_bench_n1.json
cat gpurun_out/${TAG}_bench_n1.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"march_flat|integrate_ase_owner" -s 2 -c 2 \
    -o gpurun_out/${TAG}_full -f python tools/time_cases.py ASE_medium-synth > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_full.log
ncu --set full --clock-control none --import-source on -k regex:"march_flat|integrate_seeded|integrate_scatter" -s 2 -c 2 \
    -o gpurun_out/${TAG}_seed -f python tools/time_cases.py seed_small > gpurun_out/${TAG}_ncu_seed.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_seed.log
python tools/run_configs.py > gpurun_out/${TAG}_configs.json 2> gpurun_out/${TAG}_configs.err
wc -l gpurun_out/${TAG}_configs.json; tail -2 gpurun_out/${TAG}_configs.err
