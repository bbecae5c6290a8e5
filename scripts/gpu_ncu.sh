#!/usr/bin/env bash
# One full ncu capture of the step kernel in steady state (profile_step.py: 270 settled steps, capture #262).
# Usage: scripts/gpu_ncu.sh <tag>
TAG="${1:-ncu}"
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:tds_step -s 262 -c 1 -f -o gpurun_out/${TAG}_step_full \
    python scripts/profile_step.py > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/${TAG}_step_full.ncu-rep
