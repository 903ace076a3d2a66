#!/bin/bash
# GPU visit during development: every parity test (no -x: list all failures) + one bench line with the ingest trace
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=8 ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
BOSSGPU_TRACE=1 python bench.py --no-cpu ${BENCH_ARGS:-} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"
tail -c 3000 gpurun_out/bench_quick.json; tail -8 gpurun_out/bench_quick.err
