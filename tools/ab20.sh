#!/bin/bash
# One GPU call: A/B of the variants (tools/ab_variants.py).
mkdir -p gpurun_out
timeout 900 python tools/ab_variants.py base tree joint r64 jr jr4 pf jrpf tree+ovl jr+ovl base > gpurun_out/ab20.log 2>&1
cat gpurun_out/ab20.log
