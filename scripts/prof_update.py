#!/usr/bin/env python
"""Profiling driver: device-resident strategy updates on a bench workload and nothing else (no e2e legs, no CPU arm),
so that an `ncu -k regex:... -s <skip> -c <count>` capture lands on steady-state launches.

    python scripts/prof_update.py [--workload c3] [--scale 1.0] [--updates 5] [--text]

`--text` drives the update through `process_batch_runs` (text ingest, split score pass) instead of the pre-tokenised
batch. Prints the per-kernel CUDA-event times of the last update."""
import argparse
import io
import json
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--updates", type=int, default=5)
    ap.add_argument("--text", action="store_true")
    args = ap.parse_args()
    import torch
    from boss_runs_b200 import synth
    from boss_runs_b200.hostmodel import parse_PAF
    from boss_runs_b200.runs import BossRuns
    spec = bench.workload_spec(args.workload, args.scale)
    lengths = spec["lengths"]
    names = [f"ctg{i + 1}" for i in range(len(lengths))]
    codes = bench.random_codes(lengths)
    barcodes = [f"barcode{i + 1:02d}" for i in range(spec["nb"])] if spec["nb"] > 1 else None
    run = BossRuns(contigs=dict(zip(names, codes)), ploidy=spec["ploidy"], barcodes=barcodes, bucket_threshold=5,
                   strict_upstream_asserts=False)
    run.engine.synth_coverage(seed=11, mean_depth=spec["depth"], p_ref=0.90, p_del=0.04, frac_dropout=0.02, frac_deep=0.01)
    contig_arrays = dict(zip(names, codes))
    batches = []
    for b in range(3):
        rb = synth.read_batch(contig_arrays, n_reads=spec["reads"], seed=1000 + b, mean_len=spec["mean_len"],
                              n_barcodes=spec["nb"] if spec["nb"] > 1 else 0)
        pd = parse_PAF(io.StringIO(rb.paf_text))
        for rid, recs in pd.items():
            for r in recs:
                r.barcode = rb.barcodes.get(rid) if barcodes else None
        batches.append((pd, rb.seqs))
    run.rl_dist.update({rid: recs[0].qlen for pd, _ in batches for rid, recs in pd.items()})
    if args.text:
        for i in range(args.updates):
            pd, seqs = batches[i % 3]
            run.process_batch_runs(pd, seqs)
    else:
        for pd, _ in batches:
            run.count_read_starts(pd)
        dev = []
        for pd, seqs in batches:
            d = run.pack_for_device(run.cc.convert_records(paf_dict=pd, seqs=seqs))
            dev.append({k: (torch.from_numpy(v).cuda() if isinstance(v, np.ndarray) else v) for k, v in d.items()})
        kw = dict(approx_ccl=run.rl_dist.approx_ccl, time_cost=run.rl_dist.time_cost, bucket_threshold=run.bucket_threshold)
        run.device_update(fhat_windows=run.read_starts.update_f_pointmass(), **kw)
        for i in range(args.updates):
            run.ingest_device(dev[i % 3])
            run.device_update(fhat_windows=None, **kw)
    torch.cuda.synchronize()
    print(json.dumps({"kernel_ms": run.engine.timing(), "threshold": run.last.threshold, "mirror_bytes": run.last.mirror_bytes}))


if __name__ == "__main__":
    main()
