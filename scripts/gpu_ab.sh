#!/bin/bash
# A/B on ONE box (box-to-box variation is ~10-15 %): bash scripts/gpu_ab.sh <tag> ENVVAR valA valB [more bench args]
set -u
TAG=$1; VAR=$2; A=$3; B=$4; shift 4
mkdir -p gpurun_out
for rep in 1 2; do
  for v in $A $B; do
    env $VAR=$v DDFA_BENCH_SKIP_CPU=1 timeout 600 python bench.py --steps 30 --warmup 5 "$@" > gpurun_out/${TAG}_${VAR}_${v}_r${rep}.json 2> gpurun_out/${TAG}_${VAR}_${v}_r${rep}.err
    python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_${VAR}_${v}_r${rep}.json"))
print("$VAR=$v rep $rep: %.0f graphs/s  %.4f ms/step  eager %.4f  e2e %.0f | " % (d["value"], d["ms_per_step"], d["ms_per_step_eager_instrumented"], d["e2e"]["value"]) +
      " ".join("%s %.1fus" % (l["kernel"][:12], l["avg_launch_us"]) for l in d["roofline_kernels"]))
PY
  done
done
