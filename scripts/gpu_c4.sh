#!/bin/bash
# Barcode kernel check on one B200: the C4 bench line, every GPU test that runs with more than one barcode, one ncu capture.
set -u
mkdir -p gpurun_out
timeout 200 python bench.py --workload c4 --no-cpu > gpurun_out/r02c_bench_c4.json 2> gpurun_out/r02c_bench_c4.err; echo "bench c4 exit $?"
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02c_bench_c4.json").read().splitlines() if l.startswith("{")][-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["kernel_ms"], d["checksum"])
PY
timeout 600 python -m pytest tests -m gpu -q -x -o timeout=300 -k "c4 or barcode or nb3 or nb2 or golden or split or announced" > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c_pytest_gpu.log
tail -5 gpurun_out/r02c_pytest_gpu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k_score_bin_multi' -s 8 -c 2 -f -o gpurun_out/r02c_c4 \
    python bench.py --workload c4 --no-cpu --steps 2 --warmup 3 > gpurun_out/r02c_c4_ncu.log 2>&1; echo "ncu c4 exit $?"
