#!/bin/bash
# quick GPU visit: parity tests + one bench line (+ optional trace of the ingest path)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
BOSSGPU_TRACE=1 python bench.py --no-cpu ${BENCH_ARGS:-} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"
tail -c 2600 gpurun_out/bench_quick.json; tail -8 gpurun_out/bench_quick.err
