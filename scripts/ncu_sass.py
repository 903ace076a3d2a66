#!/usr/bin/env python
"""Hot spots of one kernel from an ncu report, SASS level: share of executed warp instructions by execution count and
opcode, and the instructions where the stall samples pile up.

    python scripts/ncu_sass.py gpurun_out/x.ncu-rep k_scatter_ops [n_top]
"""
import collections
import csv
import subprocess
import sys


def main(rep, kernel, n_top=25):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kernel], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H = next(r for r in rows if "Instructions Executed" in r)
    si, ii, st, ti = (H.index(k) for k in ("Source", "Instructions Executed", "# Samples", "Thread Instructions Executed"))
    data, seen_hdr = [], 0
    for r in rows:
        if r == H:
            seen_hdr += 1
            continue
        if seen_hdr != 1 or len(r) <= max(si, ii, st, ti):
            continue
        try:
            data.append((int(r[ii] or 0), int(r[st] or 0), int(r[ti] or 0), r[si].strip()))
        except ValueError:
            pass
    tot, tots = sum(d[0] for d in data), sum(d[1] for d in data)
    print(f"{kernel}: {tot:.4g} warp instructions, {tots} stall samples, {len(data)} SASS instructions")
    by_count = collections.Counter()
    for d in data:
        by_count[d[0]] += 1
    for cnt, n in sorted(by_count.items(), key=lambda kv: -kv[0] * kv[1])[:12]:
        print(f"  executed {cnt:>10} x : {n:>5} instructions = {cnt * n / tot * 100:5.1f}% of all")
    op = collections.Counter()
    for d in data:
        t = d[3].split()
        o = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        op[o.split(".")[0]] += d[0]
    print("  opcodes:", ", ".join(f"{k} {v / tot * 100:.1f}%" for k, v in op.most_common(18)))
    print("  stall hot spots:")
    for d in sorted(data, key=lambda d: -d[1])[:n_top]:
        print(f"  {d[1] / max(tots, 1) * 100:5.1f}%  exec={d[0]:>9} thr/warp={d[2] / max(d[0], 1):4.1f}  {d[3][:110]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
