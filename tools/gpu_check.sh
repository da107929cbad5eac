#!/bin/bash
# One gpurun call: GPU tests, smoke, bench, then ncu launch list + full capture of the top kernel.
# usage: tools/gpu_check.sh <tag> [kernel-regex]
set -u
TAG=${1:-r01}
KREGEX=${2:-fused_flux}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
python bench.py --steps 1 --warmup 1 --batch 1 --no-cpu-baseline > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --batch 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:${KREGEX} -c 6 -o gpurun_out/${TAG}_prof \
    python bench.py --steps 1 --warmup 1 --batch 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out | tail -20
