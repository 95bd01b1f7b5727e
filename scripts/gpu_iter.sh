#!/bin/bash
# Development iteration on the GPU box: selectable stages, everything logged under gpurun_out/.
#   bash scripts/gpu_iter.sh <tag> <stage> [<stage> ...]
# stages: tests | tcdebug | gather | bench-simt | bench-tc | bwdbench | ncu-tc | ncu-gather | smoke
set -u
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
for STAGE in "$@"; do
  echo "=================== stage: $STAGE"
  case $STAGE in
    tests)
      timeout 900 python -m pytest tests -m gpu -q --timeout 120 -s > $OUT/${TAG}_pytest.log 2>&1
      echo "pytest exit: $?"; grep -E "passed|failed|error" $OUT/${TAG}_pytest.log | tail -n 5
      grep -E "^(FAILED|ERROR)|worst|max\|dlogit" $OUT/${TAG}_pytest.log | head -n 40 ;;
    tcdebug)
      timeout 300 python scripts/tc_debug.py > $OUT/${TAG}_tcdebug.log 2>&1; echo "tcdebug exit: $?"; cat $OUT/${TAG}_tcdebug.log | tail -n 60 ;;
    gather)
      timeout 600 python scripts/gather_bench.py > $OUT/${TAG}_gather.log 2>&1; echo "gather exit: $?"; cat $OUT/${TAG}_gather.log ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit: $?"; tail -n 4 $OUT/${TAG}_smoke.log ;;
    bench-simt)
      timeout 900 python bench.py --steps 30 --warmup 5 --engine simt > $OUT/${TAG}_bench_simt.json 2> $OUT/${TAG}_bench_simt.err; echo "exit: $?"
      cat $OUT/${TAG}_bench_simt.json; tail -n 5 $OUT/${TAG}_bench_simt.err ;;
    bench-tc)
      timeout 900 python bench.py --steps 30 --warmup 5 --engine tcgen05 > $OUT/${TAG}_bench_tc.json 2> $OUT/${TAG}_bench_tc.err; echo "exit: $?"
      cat $OUT/${TAG}_bench_tc.json; tail -n 5 $OUT/${TAG}_bench_tc.err ;;
    bench-ref)
      timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "exit: $?"
      cat $OUT/${TAG}_bench_reference.json ;;
    ncu-tc)
      DDFA_BENCH_MIN_WARMUP=1 DDFA_BENCH_SKIP_CPU=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
        --log-file $OUT/${TAG}_launches_tc.csv python bench.py --steps 2 --warmup 1 --engine tcgen05 > $OUT/${TAG}_ncu_bench_tc.log 2>&1
      echo "ncu launches exit: $?"
      DDFA_BENCH_MIN_WARMUP=1 DDFA_BENCH_SKIP_CPU=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gru_fwd3_kernel|dgrad3_kernel|wgrad_kernel|gate_bwd_image" -s 16 -c 6 \
        -f -o $OUT/${TAG}_prof_gru_tc python bench.py --steps 2 --warmup 1 --engine tcgen05 > $OUT/${TAG}_ncu_gru_tc.log 2>&1
      echo "ncu gru_tc exit: $?" ;;
    ncu-gather)
      DDFA_BENCH_MIN_WARMUP=1 DDFA_BENCH_SKIP_CPU=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather_sum -s 24 -c 3 \
        -f -o $OUT/${TAG}_prof_gather python bench.py --steps 2 --warmup 1 --engine simt > $OUT/${TAG}_ncu_gather.log 2>&1
      echo "ncu gather exit: $?" ;;
    *) echo "unknown stage $STAGE" ;;
  esac
done
ls -la $OUT | tail -n 12
