#!/usr/bin/env bash
# Round-2 GPU call aa (1 GPU): ncu capture of the D=32 / P=8 view-grid MSDA kernel at the 4K stress shape.
set -u
TAG="${1:-r02aa}"
OUT=gpurun_out
mkdir -p $OUT
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:"msda_vg" -c 2 -o $OUT/${TAG}_prof -f python scripts/prof_stress_msda.py > $OUT/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $OUT/${TAG}_ncu.log
