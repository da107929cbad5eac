#!/bin/bash
# one full ncu capture of lwa_prefix_kernel (one 0.25-degree snapshot, upsampled real data)
export FALWA_LWA_FORCE=1
CMD="python bench.py --kernels-only --steps 1 --warmup 1 --batch 1 --data upsampled"
$CMD > gpurun_out/pfx_plain.log 2>&1 || { tail -5 gpurun_out/pfx_plain.log; exit 1; }
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:lwa_prefix' -s 1 -c 1 -o gpurun_out/pfx_prof -f $CMD > gpurun_out/pfx_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/pfx_ncu.log
