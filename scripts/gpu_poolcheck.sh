#!/bin/bash
# After the host-side worker pool: ingest-heavy parity tests, then the host-bound bench lines (C2, C4) and the default line.
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_dropin.py tests/test_simulation.py -m gpu -q -x -o timeout=200 > gpurun_out/r02f_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02f_pytest.log
tail -4 gpurun_out/r02f_pytest.log
for w in c2 c4; do
  timeout 300 python bench.py --workload $w --no-cpu > gpurun_out/r02f_bench_$w.json 2> gpurun_out/r02f_bench_$w.err; echo "bench $w exit $?"
done
timeout 600 python bench.py --no-cpu > gpurun_out/r02f_bench_c3.json 2> gpurun_out/r02f_bench_c3.err; echo "bench c3 exit $?"
BOSSGPU_TRACE=1 timeout 300 python bench.py --workload c2 --no-cpu --steps 3 --warmup 3 2>&1 | grep "bossgpu\] ingest" | tail -4
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02f_bench_*.json")):
    try:
        d = json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        print(f, d["ms_per_step"], d["e2e"].get("ms_per_step"), d["e2e"].get("host_ms"), d.get("checksum"), d["e2e_from_paf_text"]["ms_per_step"])
    except Exception as e:
        print(f, "unreadable", e)
PY
