#!/bin/bash
# ncu --set full source-level captures of the tcgen05 kernels in isolation + one default bench run.
#   bash scripts/gpu_prof_tc.sh <tag>
set -u
TAG=${1:-r01i}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gru_fwd3_kernel" -s 2 -c 2 -f -o $OUT/${TAG}_prof_fwd \
  python scripts/kernel_only.py fwd > $OUT/${TAG}_ncu_fwd.log 2>&1; echo "ncu fwd exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"dgrad3_kernel|wgrad_kernel|gate_bwd" -s 3 -c 3 -f -o $OUT/${TAG}_prof_bwd \
  python scripts/kernel_only.py bwd > $OUT/${TAG}_ncu_bwd.log 2>&1; echo "ncu bwd exit $?"
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
cut -c1-400 $OUT/${TAG}_bench.json; tail -n 5 $OUT/${TAG}_bench.err
