#!/bin/bash
# One-GPU verification visit: the whole GPU suite (with durations), the default bench line, the reference arm,
# and the ncu launch list of the bench command. Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q --durations=25 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench exit $?"
tail -c 3500 gpurun_out/bench_c3.json; tail -5 gpurun_out/bench_c3.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
if [ "${NCU_LIST:-1}" = "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list exit $?"
fi
ls -la gpurun_out
