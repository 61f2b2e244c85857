#!/bin/bash
# Profiling recipe of /opt/skills/guides/B200_PROFILING.md for the fused voice-path kernel.
# Run on the GPU box:  gpurun -- 'bash profiles/capture.sh'   (outputs land in gpurun_out/)
set -u
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
mkdir -p gpurun_out
# 1) the same command must exit 0 without ncu first
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain.log; exit 1; }
# 2) every launch with its device time (cold-cache, serialised: compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 400 --csv \
    --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
# 3) the top kernel, full set, with source
ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 2 \
    -o gpurun_out/prof_fused $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
