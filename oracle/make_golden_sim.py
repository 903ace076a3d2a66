"""ORACLE / TEST INFRASTRUCTURE — golden vectors for the simulator's decision step.

Runs only in the build container (needs /root/reference). Calls the UPSTREAM functions themselves,
`BossRunsSim.make_decisions` and `BossRunsSim.filter_paf_dict` (boss/runs/simulation.py:37-135), unbound on a stub
that carries exactly the attributes they read (`contigs_filt[...].strat`, `mu`, `accept_unmapped`,
`sampler.fq_stream.read_ids`, `read_cache.mu`), on the reference's own data: the first reads of
data/BOSS_test_data/ERR3152366_10k.fq with their full-length and mu-truncated minimap2 records, against seeded random
strategies for the zymo contigs. CIGAR tags are dropped from the stored PAF text (the decision step never reads them).

    python -m oracle.make_golden_sim        # writes tests/golden/sim_decisions.npz
"""
from __future__ import annotations

import os
import sys
from pathlib import Path
from types import SimpleNamespace

REPO = Path(__file__).resolve().parent.parent
REFERENCE = Path(os.environ.get("BOSS_REFERENCE", "/root/reference"))
sys.dont_write_bytecode = True
sys.path[:0] = [str(REPO), str(REPO / "oracle" / "shims"), str(REFERENCE)]

import numpy as np  # noqa: E402

GOLDEN = REPO / "tests" / "golden"
N_READS = 900
SEED = 4


def strip_cigar(line: str) -> str:
    return "\t".join(f for f in line.rstrip("\n").split("\t") if not f.startswith("cg:Z:")) + "\n"


def read_paf(path, keep):
    out = []
    with open(path) as fh:
        for line in fh:
            if line.split("\t", 1)[0] in keep:
                out.append(strip_cigar(line))
    return "".join(out)


def strategies(lengths: dict, nb: int, seed: int) -> dict:
    """Seeded random strategies: long accept/reject runs, like real masks."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, L in lengths.items():
        rows = L // 100
        runs = rng.integers(1, 400, size=rows // 50 + 2)
        vals = rng.random(size=runs.shape[0]) < 0.55
        col = np.repeat(vals, runs)[:rows]
        arr = np.stack([np.roll(col, int(rng.integers(rows))) for _ in range(2 * nb)], axis=1).reshape(rows, 2, nb)
        out[name] = arr
    return out


def main():
    import logging
    logging.disable(logging.CRITICAL)
    from boss.runs.simulation import BossRunsSim

    data = REFERENCE / "data" / "BOSS_test_data"
    lengths, name, n = {}, None, 0
    for line in open(data / "zymo.fa"):
        if line.startswith(">"):
            if name is not None:
                lengths[name] = n
            name, n = line[1:].split()[0], 0
        else:
            n += len(line.strip())
    lengths[name] = n
    tracked = {k: v for k, v in lengths.items() if v >= 100_000}
    # one tracked contig is left WITHOUT a strategy: reads mapping there are rejected by the KeyError branch
    missing = "NZ_CP041013.1"       # 35 of the 900 reads map there
    rids, rlen = [], []
    with open(data / "ERR3152366_10k.fq") as fh:
        while len(rids) < N_READS:
            head = fh.readline()
            if not head:
                break
            seq = fh.readline().strip()
            fh.readline(); fh.readline()
            rids.append(head[1:].split()[0]); rlen.append(len(seq))
    keep = set(rids)
    paf_full = read_paf(data / "ERR3152366_10k.paf", keep)
    paf_trunc = read_paf(data / "ERR3152366_10k_trunc.paf", keep)
    seqs = {r: "A" * n for r, n in zip(rids, rlen)}
    out = {"rids": np.array(rids), "read_len": np.array(rlen, dtype=np.int64), "paf_full": np.frombuffer(paf_full.encode(), np.uint8),
           "paf_trunc": np.frombuffer(paf_trunc.encode(), np.uint8), "contigs": np.array(list(tracked)),
           "contig_len": np.array(list(tracked.values()), dtype=np.int64), "missing": np.array(missing), "seed": np.int64(SEED)}
    for nb, accept_unmapped in ((1, True), (1, False), (3, True)):
        strat = strategies(tracked, nb, SEED + nb)
        stub = SimpleNamespace(contigs_filt={k: SimpleNamespace(strat=v) for k, v in strat.items() if k != missing}, mu=400,
                               accept_unmapped=accept_unmapped, read_cache=SimpleNamespace(mu=400),
                               sampler=SimpleNamespace(fq_stream=SimpleNamespace(read_ids=set(rids))))
        rng = np.random.default_rng(100 + nb)
        barcodes = {r: int(rng.integers(nb)) for r in rids}
        paf_dict, reads_decision, n_mapped, n_unmapped, n_acc, n_rej = BossRunsSim.make_decisions(
            stub, seqs=seqs, paf_full=paf_full, paf_trunc=paf_trunc, barcodes=barcodes)
        acc = BossRunsSim.filter_paf_dict(stub, paf_dict=paf_dict)
        tag = f"nb{nb}_{'acc' if accept_unmapped else 'rej'}_"
        keys = list(paf_dict)
        recs = [paf_dict[k][0] for k in keys]
        out[tag + "keys"] = np.array(keys)
        out[tag + "n_recs"] = np.array([len(paf_dict[k]) for k in keys], dtype=np.int64)
        out[tag + "rec"] = np.array([(r.qlen, r.qstart, r.qend, r.tstart, r.tend, r.rev, r.mapq, r.align_score) for r in recs], dtype=np.int64)
        out[tag + "rec_tname"] = np.array([str(r.tname) for r in recs])
        out[tag + "rec_barcode"] = np.array([-1 if r.barcode is None else int(r.barcode) for r in recs], dtype=np.int64)
        out[tag + "counts"] = np.array([n_mapped, n_unmapped, n_acc, n_rej], dtype=np.int64)
        out[tag + "decision_len"] = np.array([len(reads_decision[r]) for r in rids], dtype=np.int64)
        out[tag + "accepted_keys"] = np.array(list(acc))
        out[tag + "barcodes"] = np.array([barcodes[r] for r in rids], dtype=np.int64)
        print(tag, "mapped/unmapped/accepted/rejected", n_mapped, n_unmapped, n_acc, n_rej, "accepted in paf_dict", len(acc))
    GOLDEN.mkdir(parents=True, exist_ok=True)
    path = GOLDEN / "sim_decisions.npz"
    np.savez_compressed(path, **out)
    print(path.name, f"{path.stat().st_size / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
