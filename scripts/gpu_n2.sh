#!/bin/bash
# N-GPU bench lines on one box (run under `gpurun --gpus N`):  bash scripts/gpu_n2.sh <tag> <N>
# exchange A/B on one box: NCCL split (default) / fused peer-memory kernel / NCCL single call — C1 per GPU, then C0 (where the
# exchange is 4x larger relative to the step)
set -u
TAG=$1; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
run() {  # name, env, args
  env $2 DDFA_BENCH_SKIP_CPU=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-variable --no-secondary ${QUICK:-} $3 > $OUT/${TAG}_$1.json 2> $OUT/${TAG}_$1.err
  echo "$1 exit $?"; python - <<PY
import json
try:
    d = json.loads([l for l in open("$OUT/${TAG}_$1.json").read().splitlines() if l.startswith("{")][-1])
    print("$1: %.0f graphs/s %.4f ms/step e2e %.0f | dp_parity %s | %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("dp_parity") and {k: d["dp_parity"][k] for k in ("max_abs_loss_diff", "max_abs_param_diff")}, d.get("allreduce")))
except Exception as e:
    print("$1: no line", e); print(open("$OUT/${TAG}_$1.err").read()[-1500:])
PY
}
if [ "${ONLY:-}" = "c0" ]; then
  run c0_split "DDFA_AR_OVERLAP=1" "--graphs 256"
  run c0_p2p "DDFA_EXCHANGE=p2p" "--graphs 256"
  run c0_single "DDFA_AR_OVERLAP=0" "--graphs 256"
  run c1_p2p "DDFA_EXCHANGE=p2p" ""
elif [ "${ONLY:-}" = "ab" ]; then
  run c1_split "DDFA_AR_OVERLAP=1" ""
  run c1_p2p "DDFA_EXCHANGE=p2p" ""
  run c0_split "DDFA_AR_OVERLAP=1" "--graphs 256"
  run c0_p2p "DDFA_EXCHANGE=p2p" "--graphs 256"
else
  run c1_split "DDFA_AR_OVERLAP=1" ""
  run c1_p2p "DDFA_EXCHANGE=p2p" ""
  run c0_split "DDFA_AR_OVERLAP=1" "--graphs 256"
  run c0_p2p "DDFA_EXCHANGE=p2p" "--graphs 256"
  run c0_single "DDFA_AR_OVERLAP=0" "--graphs 256"
fi
