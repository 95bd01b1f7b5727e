#!/bin/bash
# One gpurun call: GPU tests, smoke, bench (ours + reference arm), ncu launch list and full captures.
# Usage (from the repo root on the GPU box):  bash scripts/gpu_check.sh [tag] [engine]
set -u
TAG=${1:-r01}
ENGINE=${2:-simt}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > $OUT/${TAG}_gpu.txt 2>&1
echo "== pytest -m gpu" | tee $OUT/${TAG}_pytest.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 300 -s >> $OUT/${TAG}_pytest.log 2>&1
echo "pytest exit: $?" | tee -a $OUT/${TAG}_pytest.log
tail -n 25 $OUT/${TAG}_pytest.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit: $?" | tee -a $OUT/${TAG}_smoke.log
tail -n 5 $OUT/${TAG}_smoke.log
echo "== bench ours ($ENGINE)"
timeout 900 python bench.py --gpus 1 --steps 30 --warmup 5 --engine $ENGINE > $OUT/${TAG}_bench_${ENGINE}.json 2> $OUT/${TAG}_bench_${ENGINE}.err; echo "bench exit: $?"
cat $OUT/${TAG}_bench_${ENGINE}.json; tail -n 5 $OUT/${TAG}_bench_${ENGINE}.err
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 10 --warmup 2 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; echo "ref exit: $?"
cat $OUT/${TAG}_bench_reference.json
echo "== ncu launch list"
DDFA_BENCH_MIN_WARMUP=1 DDFA_BENCH_SKIP_CPU=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file $OUT/${TAG}_launches_${ENGINE}.csv python bench.py --steps 2 --warmup 1 --engine $ENGINE > $OUT/${TAG}_ncu_bench.log 2>&1
echo "ncu launches exit: $?"; wc -l $OUT/${TAG}_launches_${ENGINE}.csv
echo "== ncu full: gather"
DDFA_BENCH_MIN_WARMUP=1 DDFA_BENCH_SKIP_CPU=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather_sum -s 24 -c 3 \
  -f -o $OUT/${TAG}_prof_gather python bench.py --steps 2 --warmup 1 --engine $ENGINE > $OUT/${TAG}_ncu_gather.log 2>&1
echo "ncu gather exit: $?"
echo "== ncu full: GRU GEMM"
DDFA_BENCH_MIN_WARMUP=1 DDFA_BENCH_SKIP_CPU=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"sgemm_kernel|gru_tc" -s 40 -c 3 \
  -f -o $OUT/${TAG}_prof_gru python bench.py --steps 2 --warmup 1 --engine $ENGINE > $OUT/${TAG}_ncu_gru.log 2>&1
echo "ncu gru exit: $?"
ls -la $OUT | tail -n 20
