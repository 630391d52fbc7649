#!/bin/bash
# The reference's own driver (unmodified sources + the registration patch of INTEGRATION.md) with
# every method in one process; needs oracle/_ref built (`make -C oracle legacy`, in the container
# that has /root/reference).  Run under gpurun from the repo root (any number of GPUs: the
# b200-multigpu method uses all of them).
python tools/make_dat.py .bench_tmp ase_small seed_small ase_medium_synth
B=./oracle/_ref/CreateImageB200_legacy
nvidia-smi -L | wc -l
$B -iterations=5 -methods=cpu,threads,Cuda,b200,b200-direct,b200-multigpu .bench_tmp/ASE_small.dat 2>&1 | grep -v "^$" | tail -10
$B -iterations=5 -methods=threads,Cuda,b200,b200-direct,b200-multigpu .bench_tmp/ASE_medium_synth.dat 2>&1 | grep -v "^$" | tail -10
$B -iterations=3 -methods=${SEED_METHODS:-cpu,Cuda,b200,b200-direct,b200-multigpu} .bench_tmp/seed_small.dat 2>&1 | grep -v "^$" | tail -13
