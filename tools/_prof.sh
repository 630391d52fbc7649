# scratch: untimed run first, then one ncu capture of the ASE_medium-synth kernels
python tools/time_cases.py ASE_medium-synth || exit 1
ncu --set full --clock-control none --import-source on -k regex:"march_flat|integrate_ase_owner" -s 2 -c 2 -o gpurun_out/prof_r1b -f python tools/time_cases.py ASE_medium-synth > gpurun_out/prof_r1b.log 2>&1
tail -2 gpurun_out/prof_r1b.log
