#!/bin/bash
# one full ncu capture of selected kernels inside the bench command; usage: gpu_ncu.sh '<kernel regex>' <skip> <count> <out>
set -u
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$1" -s ${2:-6} -c ${3:-6} -f -o gpurun_out/${4:-prof} \
    python bench.py --steps 2 --warmup 3 --no-cpu ${BENCH_ARGS:-} > gpurun_out/${4:-prof}.log 2>&1; echo "ncu exit $?"
tail -3 gpurun_out/${4:-prof}.log | cut -c1-600
