python tools/time_cases.py ASE_medium-synth || exit 1
ncu --set full --clock-control none --import-source on -k regex:"integrate_ase_packed" -s 1 -c 1 -o gpurun_out/prof_r1e -f python tools/time_cases.py ASE_medium-synth > gpurun_out/prof_r1e.log 2>&1
tail -2 gpurun_out/prof_r1e.log
