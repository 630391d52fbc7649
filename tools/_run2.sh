python tools/make_dat.py .bench_tmp ase_small seed_small
./oracle/_ref/CreateImageB200 -iterations=3 -methods=cpu,threads,b200,b200-multigpu .bench_tmp/ASE_small.dat 2>&1 | tail -16
./oracle/_ref/CreateImage_b200 -iterations=3 -methods=b200 .bench_tmp/ASE_small.dat 2>&1 | tail -6
./oracle/_ref/CreateImageB200 -iterations=1 -methods=cpu,b200 .bench_tmp/seed_small.dat 2>&1 | tail -10
