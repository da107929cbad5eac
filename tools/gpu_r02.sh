#!/bin/bash
# One gpurun call of round 2: bench line, ncu launch list, full capture of one kernel.
# usage: tools/gpu_r02.sh <tag> <kernel-regex> [extra bench args]
set -u
TAG=$1; KREGEX=$2; shift 2
mkdir -p gpurun_out
python bench.py --no-cpu-baseline "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "roofline", d["roofline"]["kernel"], d["roofline"]["frac"], "pipeline_frac", d["roofline"]["pipeline_frac"])
for k,v in d["kernels"].items(): print("  %-34s %7.3f ms/step  share %.3f  launches %.1f"%(k, v["ms_per_step"], v["share"], v["launches_per_step"]))
PY
python bench.py --steps 1 --warmup 1 --batch 1 --no-cpu-baseline > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -c 3 -o gpurun_out/${TAG}_prof -f \
    python bench.py --steps 1 --warmup 1 --batch 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
