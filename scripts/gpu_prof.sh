#!/bin/bash
# ncu captures of the update's kernels at steady state; usage: gpu_prof.sh '<kernel regex>' <skip> <count> <name> [prof_update args]
set -u
mkdir -p gpurun_out
re="$1"; skip="$2"; cnt="$3"; name="$4"; shift 4
ncu --set full --clock-control none --import-source on -k regex:"$re" -s "$skip" -c "$cnt" -f -o gpurun_out/$name \
    python scripts/prof_update.py "$@" > gpurun_out/$name.log 2>&1; echo "ncu exit $?"
tail -2 gpurun_out/$name.log | cut -c1-800
