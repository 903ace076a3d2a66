#!/bin/bash
# tests + bench lines of every workload (+ optional ncu capture); everything lands in gpurun_out/
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --durations=8 ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for w in ${WORKLOADS:-c3}; do
  extra=""; [ "$w" != "c3" ] && extra="--no-cpu"
  python bench.py --workload $w $extra ${BENCH_ARGS:-} > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w exit $?"
  tail -c 2500 gpurun_out/bench_$w.json; tail -3 gpurun_out/bench_$w.err
done
