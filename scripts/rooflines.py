#!/usr/bin/env python
"""Per-kernel roofline fractions of one bench line: CUDA-event time of every update kernel (bench.py `kernel_ms`) against its
algorithmic bytes and the measured HBM peak the line itself reports.

    python scripts/rooflines.py profiles/r02d_bench_c3_n1.json > profiles/r02d_rooflines_c3.json
"""
import json
import sys


def main(path):
    d = json.loads([l for l in open(path).read().splitlines() if l.startswith("{")][-1])
    sites, nb = d["config"]["sites"], d["config"]["barcodes"]
    rows = sites / 100.0 * nb                         # merged rows ~ bins (one extra row per contig is noise here)
    reads = d["config"]["reads_per_batch"]
    positions = reads * 10_000                        # aligned reference positions of one batch (mean read 10 kb)
    peak = d["roofline"]["peak"]
    alg = {
        "score_bin": (d["roofline"]["algorithmic_bytes_per_launch"], "11 B/site (10 B counters + 1 B reference base) + 8 B per bin written"),
        "smooth": (rows * (8 + 16), "8 B/bin read + 16 B/row (forward, reverse benefit) written"),
        "hist": (rows * (16 + 2), "16 B/row read + 2 code bytes/row written"),
        "distribute": (rows * (2 + 2), "2 code bytes/row read + 2 mask bytes/row compared (written where changed)"),
        "scatter": (positions * 5, "1 B base + 2 B counter read + 2 B counter written per aligned position (SURVEY 8d)"),
    }
    out = {"source": path, "peak_gbs": peak, "peak_source": d["roofline"]["peak_source"], "ms_per_update": d["ms_per_step"], "kernels": {}}
    total_bytes = 0.0
    for k, (b, what) in alg.items():
        ms = d["kernel_ms"][k]
        gbs = b / (ms * 1e-3) / 1e9
        total_bytes += b
        out["kernels"][k] = {"ms": round(ms, 4), "algorithmic_bytes": int(b), "basis": what, "achieved_gbs": round(gbs, 1), "frac": round(gbs / peak, 3)}
    out["kernels"]["threshold"] = {"ms": round(d["kernel_ms"]["threshold"], 4), "note": "one CTA, 26 KB of input: latency, no roofline"}
    whole = total_bytes / (d["ms_per_step"] * 1e-3) / 1e9
    out["whole_update"] = {"algorithmic_bytes": int(total_bytes), "achieved_gbs": round(whole, 1), "frac": round(whole / peak, 3)}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
