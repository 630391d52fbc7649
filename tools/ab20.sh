#!/bin/bash
# One GPU call: the -m gpu suite on the in-tree build, with the overlapped launch, and on the
# candidate build; then the A/B of the variants (tools/ab_variants.py).
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/ab20_t_tree.log 2>&1; echo "tree rc=$?" > gpurun_out/ab20_rc.log
RTB200_OVERLAP=1 timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/ab20_t_ovl.log 2>&1; echo "ovl rc=$?" >> gpurun_out/ab20_rc.log
RTB200_LIB=$PWD/.variants/jr.so timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/ab20_t_jr.log 2>&1; echo "jr rc=$?" >> gpurun_out/ab20_rc.log
timeout 900 python tools/ab_variants.py base tree joint r64 jr jr4 pf jrpf tree+ovl jr+ovl base > gpurun_out/ab20.log 2>&1
cat gpurun_out/ab20_rc.log; tail -3 gpurun_out/ab20_t_tree.log gpurun_out/ab20_t_ovl.log gpurun_out/ab20_t_jr.log; cat gpurun_out/ab20.log
