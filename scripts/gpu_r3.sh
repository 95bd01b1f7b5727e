#!/bin/bash
# Round-3 GPU iteration: selectable stages, everything logged under gpurun_out/.
#   bash scripts/gpu_r3.sh <tag> <stage> [<stage> ...]
# stages: tests | testsx (stop at first failure) | bench | bench-c0 | bench-ref | ncu-list | ncu-full | ncu-gather | smoke | gatherab | kernel
set -u
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
PROF_ENV="DDFA_BENCH_MIN_WARMUP=1 DDFA_BENCH_SKIP_CPU=1 DDFA_BENCH_MIN_TIMED_MS=0"
for STAGE in "$@"; do
  echo "=================== stage: $STAGE"
  case $STAGE in
    tests|testsx)
      X=""; [ $STAGE = testsx ] && X="-x"
      timeout 1800 python -m pytest tests -m gpu -q --timeout 600 -s $X > $OUT/${TAG}_pytest.log 2>&1
      echo "pytest exit: $?"; grep -E "passed|failed|error" $OUT/${TAG}_pytest.log | tail -n 5
      grep -E "^(FAILED|ERROR)|worst|max\|d|agreement|F1|adam x3|C1 variable|gradient worst" $OUT/${TAG}_pytest.log | head -n 60 ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit: $?"; tail -n 4 $OUT/${TAG}_smoke.log ;;
    bench)
      timeout 1200 python bench.py > $OUT/${TAG}_bench_c1.json 2> $OUT/${TAG}_bench_c1.err; echo "exit: $?"
      cat $OUT/${TAG}_bench_c1.json; tail -n 5 $OUT/${TAG}_bench_c1.err ;;
    bench-c0)
      timeout 900 python bench.py --graphs 256 --no-variable > $OUT/${TAG}_bench_c0.json 2> $OUT/${TAG}_bench_c0.err; echo "exit: $?"
      cat $OUT/${TAG}_bench_c0.json; tail -n 5 $OUT/${TAG}_bench_c0.err ;;
    bench-ref)
      timeout 900 python bench.py --impl reference --steps 6 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "exit: $?"
      cat $OUT/${TAG}_bench_reference.json ;;
    ncu-list)
      env $PROF_ENV timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv \
        --log-file $OUT/${TAG}_launches_c1.csv python bench.py --steps 2 --warmup 1 --no-secondary --no-variable > $OUT/${TAG}_ncu_list.log 2>&1
      echo "ncu launches exit: $?" ;;
    ncu-full)
      env $PROF_ENV timeout 1500 ncu --set full --clock-control none --import-source on \
        -k regex:"gru_fwd3_kernel|dgrad3_kernel|wgrad_kernel|gate_bwd_image|gather_sum_image|readout|embed_concat" -s 60 -c 48 \
        -f -o $OUT/${TAG}_prof_c1 python bench.py --steps 2 --warmup 1 --no-secondary --no-variable > $OUT/${TAG}_ncu_full.log 2>&1
      echo "ncu full exit: $?" ;;
    ab-packed)    # whole-step A/B on this box: round-1 saved state (fp32 h_t + four fp32 gate planes) vs packed state
      bash scripts/gpu_ab.sh ${TAG} DDFA_PACKED_STATE 0 1 --no-secondary --no-variable 2>&1 | tee $OUT/${TAG}_ab_packed_c1.log
      bash scripts/gpu_ab.sh ${TAG}c0 DDFA_PACKED_STATE 0 1 --graphs 256 --no-variable 2>&1 | tee $OUT/${TAG}_ab_packed_c0.log ;;
    pairtest)     # the CTA-pair forward kernel alone, short timeout (first runs of a new synchronisation protocol)
      timeout 240 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 200 -s -x -k "cta_pair" > $OUT/${TAG}_pytest_pair.log 2>&1
      echo "pair pytest exit: $?"; tail -n 12 $OUT/${TAG}_pytest_pair.log | cut -c1-300 ;;
    ab-pair)
      bash scripts/gpu_ab.sh ${TAG} DDFA_FWD_PAIR 0 1 --no-secondary --no-variable 2>&1 | tee $OUT/${TAG}_ab_pair_c1.log
      bash scripts/gpu_ab.sh ${TAG}c0 DDFA_FWD_PAIR 0 1 --graphs 256 --no-variable 2>&1 | tee $OUT/${TAG}_ab_pair_c0.log ;;
    kerneltests)
      timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 300 -s -x > $OUT/${TAG}_pytest_kernels.log 2>&1
      echo "pytest exit: $?"; tail -n 15 $OUT/${TAG}_pytest_kernels.log | cut -c1-300 ;;
    gatherab)
      timeout 900 python scripts/gather_bench.py > $OUT/${TAG}_gather_ab.log 2>&1; echo "gather exit: $?"; cat $OUT/${TAG}_gather_ab.log ;;
    *) echo "unknown stage $STAGE" ;;
  esac
done
ls -la $OUT | tail -n 12
