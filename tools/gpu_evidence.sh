#!/bin/bash
# One gpurun call that collects the evidence kept under profiles/: bench lines of both arms and of configs 4 / 5, the
# FP64 micro-benchmark, the ncu launch list and one full capture of every hot kernel.   usage: tools/gpu_evidence.sh <tag>
set -u
T=${1:-r02}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/${T}_gpu.txt 2>&1
python bench.py --cpu-extra > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
python bench.py --data upsampled --no-cpu-baseline > $O/${T}_bench_upsampled.json 2>> $O/${T}_bench.err; echo "upsampled rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_reference.json 2>> $O/${T}_bench.err; echo "reference rc=$?"
python bench.py --config 4 --steps 2 --warmup 1 > $O/${T}_config4.json 2>> $O/${T}_bench.err; echo "config4 rc=$?"
python bench.py --config 4 --impl reference --steps 1 --warmup 1 >> $O/${T}_config4.json 2>> $O/${T}_bench.err
python bench.py --config 5 > $O/${T}_config5.json 2>> $O/${T}_bench.err; echo "config5 rc=$?"
python bench.py --config 5 --impl reference >> $O/${T}_config5.json 2>> $O/${T}_bench.err
python tools/fp64_peak.py > $O/${T}_fp64_peak.json 2>> $O/${T}_bench.err; echo "fp64 rc=$?"
python tools/membw.py > $O/${T}_membw.txt 2>&1
CMD="python bench.py --kernels-only --steps 1 --warmup 1 --batch 1"
$CMD > $O/${T}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_ncu_launches.csv $CMD > $O/${T}_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:interp_qgpv|lwa_window|flux_column|area_hist3|zonal_stats|theta_rowmean' \
    -s 10 -c 8 -o $O/${T}_prof -f $CMD > $O/${T}_ncu_full.log 2>&1
echo "ncu full rc=$?"
ls -la $O | grep ${T}_
