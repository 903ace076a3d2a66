#!/bin/bash
# Round-end evidence on one B200: quick parity subset, bench lines of every workload, the reference arm, the ncu launch list of
# the bench command and one full ncu capture of every update kernel at steady state. Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_dropin.py tests/test_aeons.py -m gpu -q -o timeout=150 > gpurun_out/pytest_gpu_quick.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_quick.log
tail -4 gpurun_out/pytest_gpu_quick.log
timeout 600 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench c3 exit $?"
for w in c2 c4 c5; do
  timeout 300 python bench.py --workload $w --no-cpu > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "bench $w exit $?"
done
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
# launch list of the bench command (shares, not absolutes)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list exit $?"
# one full capture of every kernel of one steady-state update (device-resident batch)
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'k_score_bin_tma|k_smooth|k_hist|k_distribute|k_scatter_ops|k_op_prefix|k_check_bases|k_threshold|k_tile_reduce|k_buckets' -s 30 -c 10 -f -o gpurun_out/r02_update \
    python scripts/prof_update.py --updates 5 > gpurun_out/r02_update.log 2>&1; echo "ncu full exit $?"
# ... and of the text ingest + split pass (process_batch_runs)
timeout 600 ncu --set full --clock-control none \
    -k regex:'k_tokenize|k_op_prefix|k_scatter_ops|k_score_bin_tma|k_mark' -s 24 -c 8 -f -o gpurun_out/r02_ingest \
    python scripts/prof_update.py --updates 5 --text > gpurun_out/r02_ingest.log 2>&1; echo "ncu ingest exit $?"
ls -la gpurun_out | tail -20
