#!/bin/bash
# Profiling visit (one GPU): the default bench line, the reference arm, the ncu launch list of the same command and one
# full capture of the update's kernels. Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench exit $?"
tail -c 3800 gpurun_out/bench_c3.json; tail -5 gpurun_out/bench_c3.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'k_score_bin_tma|k_smooth|k_hist|k_distribute|k_scatter|k_mark_tiles|k_threshold' -s 40 -c 16 -f -o gpurun_out/prof_update \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu_full.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out
