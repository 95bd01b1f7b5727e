#!/bin/bash
# ncu --set full captures of the four tcgen05-engine kernels in isolation (scripts/kernel_only.py) -> gpurun_out/<tag>_prof_{fwd,bwd}.ncu-rep
set -u
TAG=${1:-r02o}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gru_fwd3_kernel" -s 2 -c 2 -f -o $OUT/${TAG}_prof_fwd \
  python scripts/kernel_only.py fwd > $OUT/${TAG}_ncu_fwd.log 2>&1; echo "ncu fwd exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"dgrad3_kernel|wgrad_kernel|gate_bwd" -s 3 -c 3 -f -o $OUT/${TAG}_prof_bwd \
  python scripts/kernel_only.py bwd > $OUT/${TAG}_ncu_bwd.log 2>&1; echo "ncu bwd exit $?"
