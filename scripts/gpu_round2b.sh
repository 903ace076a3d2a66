#!/bin/bash
# Second evidence visit of round 2 (one B200): the rewritten barcode kernel first (time-boxed), then the whole GPU suite, the
# bench lines again and one full ncu capture of the barcode kernel. Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 200 python bench.py --workload c4 --no-cpu > gpurun_out/r02b_bench_c4.json 2> gpurun_out/r02b_bench_c4.err; rc=$?; echo "bench c4 exit $rc"
tail -c 700 gpurun_out/r02b_bench_c4.json
if [ $rc -ne 0 ]; then
  tail -5 gpurun_out/r02b_bench_c4.err
  echo "barcode kernel failed its bench: the suite runs on the two-pass path"; export BOSSGPU_NO_FUSED_BARCODES=1
fi
timeout 1100 python -m pytest tests -m gpu -q -x -o timeout=300 > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02b_pytest_gpu.log
tail -6 gpurun_out/r02b_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02b_bench_c3.json 2> gpurun_out/r02b_bench_c3.err; echo "bench c3 exit $?"
for w in c5 c2; do
  timeout 300 python bench.py --workload $w --no-cpu > gpurun_out/r02b_bench_$w.json 2> gpurun_out/r02b_bench_$w.err; echo "bench $w exit $?"
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_score_bin_multi|k_threshold' -s 8 -c 4 -f -o gpurun_out/r02b_c4 \
    python bench.py --workload c4 --no-cpu --steps 2 --warmup 3 > gpurun_out/r02b_c4_ncu.log 2>&1; echo "ncu c4 exit $?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02b_bench_*.json")):
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["kernel_ms"], d["e2e"].get("host_ms"))
    except Exception as e:
        print(f, "unreadable", e)
PY
