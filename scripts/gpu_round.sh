#!/bin/bash
# One GPU visit: parity tests, the bench line, the ncu launch list and one full capture of the hot kernels.
# Run under gpurun from the repo root; everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench exit $?"
tail -c 3000 gpurun_out/bench_c3.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
# launch list of the same command (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list exit $?"
# one full capture of the kernels of an update (2 updates' worth after the warm-up ones)
ncu --set full --clock-control none --import-source on \
    -k regex:'k_score_bin_tma|k_smooth|k_hist|k_distribute|k_scatter' -s 15 -c 10 -f -o gpurun_out/prof_update \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out
