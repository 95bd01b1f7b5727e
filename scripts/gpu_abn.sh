#!/bin/bash
# N-way A/B on one box: bash scripts/gpu_abn.sh <tag> ENVVAR v1 v2 v3 ...
set -u
TAG=$1; VAR=$2; shift 2
mkdir -p gpurun_out
for v in "$@"; do
  env $VAR=$v DDFA_BENCH_SKIP_CPU=1 timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/${TAG}_${VAR}_${v}.json 2> gpurun_out/${TAG}_${VAR}_${v}.err
  python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_${VAR}_${v}.json"))
print("$VAR=$v: %.0f graphs/s  %.4f ms/step  e2e %.0f | " % (d["value"], d["ms_per_step"], d["e2e"]["value"]) +
      " ".join("%s %.1fus" % (l["kernel"][:12], l["avg_launch_us"]) for l in d["roofline_kernels"]))
PY
done
