#!/bin/bash
# Final evidence visit of round 2 (one B200): whole GPU suite, bench lines of every workload, the reference arm, the ncu launch
# list of the bench command and a full ncu capture of the update kernels. Everything lands in gpurun_out/ with the r02d_ prefix.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1100 python -m pytest tests -m gpu -q -x -o timeout=300 > gpurun_out/r02d_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02d_pytest_gpu.log
tail -4 gpurun_out/r02d_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02d_bench_c3.json 2> gpurun_out/r02d_bench_c3.err; echo "bench c3 exit $?"
for w in c2 c4 c5; do
  timeout 300 python bench.py --workload $w --no-cpu > gpurun_out/r02d_bench_$w.json 2> gpurun_out/r02d_bench_$w.err; echo "bench $w exit $?"
done
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02d_bench_ref.json 2> gpurun_out/r02d_bench_ref.err; echo "ref exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02d_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02d_bench_under_ncu.log 2>&1; echo "ncu list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'k_score_bin_tma|k_smooth|k_hist|k_distribute|k_scatter_ops|k_op_prefix|k_check_bases|k_threshold|k_tile_reduce|k_buckets' -s 30 -c 10 -f -o gpurun_out/r02d_update \
    python scripts/prof_update.py --updates 5 > gpurun_out/r02d_update.log 2>&1; echo "ncu full exit $?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02d_bench_*.json")):
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        print(f, d["ms_per_step"], d["e2e"].get("ms_per_step"), (d.get("roofline") or {}).get("frac"), d.get("kernel_ms"), d["e2e"].get("host_ms"), d.get("checksum"))
    except Exception as e:
        print(f, "unreadable", e)
PY
