python tools/make_dat.py .bench_tmp ase_small seed_small ase_medium_synth
./oracle/_ref/CreateImageB200_legacy -iterations=5 -methods=cpu,threads,Cuda,b200,b200-direct .bench_tmp/ASE_small.dat 2>&1 | grep -v "^$" | tail -9
./oracle/_ref/CreateImageB200_legacy -iterations=3 -methods=threads,Cuda,b200,b200-direct .bench_tmp/ASE_medium_synth.dat 2>&1 | grep -v "^$" | tail -9
./oracle/_ref/CreateImageB200_legacy -iterations=2 -methods=cpu,Cuda,b200,b200-direct .bench_tmp/seed_small.dat 2>&1 | grep -v "^$" | tail -12
