"""ORACLE / TEST INFRASTRUCTURE — golden vectors for the simulator's decision step.

Runs only in the build container (needs /root/reference). Calls the UPSTREAM functions themselves,
`BossRunsSim.make_decisions` and `BossRunsSim.filter_paf_dict` (boss/runs/simulation.py:37-135), unbound on a stub
that carries exactly the attributes they read (`contigs_filt[...].strat`, `mu`, `accept_unmapped`,
`sampler.fq_stream.read_ids`, `read_cache.mu`), on the reference's own data: the first reads of
data/BOSS_test_data/ERR3152366_10k.fq with their full-length and mu-truncated minimap2 records, against seeded random
strategies for the zymo contigs. CIGAR tags are dropped from the stored PAF text (the decision step never reads them).

A second fixture pins the whole simulated flow (simulation.py:139-190 minus sampler and read cache): on the inputs of
tests/golden/case_real_zymo.npz (real reads + their full-length records) plus the same reads' mu-truncated records,
the reference's `BossRuns` is driven batch by batch — decisions from its current strategies, read lengths and read
starts from the accepted reads, coverage from every record (Q12: rejected reverse-strand reads are sliced from the
whole read with qlen = 400 coordinates) — and its strategies, thresholds and counts are recorded.

    python -m oracle.make_golden_sim        # writes tests/golden/sim_decisions.npz and tests/golden/sim_flow_zymo.npz
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


def trunc_text_for(rids) -> str:
    keep, out = set(rids), []
    with open(REFERENCE / "data" / "BOSS_test_data" / "ERR3152366_10k_trunc.paf") as fh:
        for line in fh:
            if line.split("\t", 1)[0] in keep:
                out.append(line)
    return "".join(out)


def main_flow():
    import hashlib
    import io
    import logging
    import tempfile
    logging.disable(logging.CRITICAL)
    from oracle import make_golden as mg
    from boss.config import Config
    from boss.runs.core import BossRuns
    from boss.runs.simulation import BossRunsSim

    spec, names, seqs, kinds, batches = mg.real_case_inputs()
    out = {"n_batches": np.int64(len(batches))}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        work = Path(td)
        fa = work / "ref.fa"
        with open(fa, "w") as fh:
            for n in names:
                fh.write(f">{n}\n{seqs[n]}\n")
        (work / "ref.mmi").touch()
        os.chdir(work)
        try:
            args = Config().args
            args.general.name = "golden"
            args.general.ref = str(fa)
            args.general.mmi = str(work / "ref.mmi")
            args.optional.ploidy = 1
            args.optional.bucket_threshold = 0
            exp = BossRuns(args)
            exp.init()
            all_ids = set(r for rb in batches for r in rb.seqs)
            stub = SimpleNamespace(contigs_filt=exp.contigs_filt, mu=400, accept_unmapped=False, read_cache=SimpleNamespace(mu=400),
                                   sampler=SimpleNamespace(fq_stream=SimpleNamespace(read_ids=all_ids)))
            for bi, rb in enumerate(batches):
                paf_t = trunc_text_for(rb.seqs)
                barcodes = {r: 0 for r in rb.seqs}
                paf_dict, reads_decision, n_mapped, n_unmapped, n_acc, n_rej = BossRunsSim.make_decisions(
                    stub, seqs=rb.seqs, paf_full=rb.paf_text, paf_trunc=paf_t, barcodes=barcodes)
                acc = BossRunsSim.filter_paf_dict(stub, paf_dict=paf_dict)
                exp.rl_dist.update(read_lengths={n: r[0].qlen for n, r in acc.items()})
                quals = {rid: "5" * len(s) for rid, s in rb.seqs.items()}
                inc = exp.cc.convert_records(paf_dict=paf_dict, seqs=rb.seqs, quals=quals, barcodes=barcodes)
                exp._effect_increments(increments=inc)
                exp.read_starts.count_read_starts(paf_dict=acc)
                exp.update_wrapper()
                p = f"b{bi}_"
                out[p + "paf_trunc"] = np.frombuffer(paf_t.encode(), np.uint8)
                out[p + "counts"] = np.array([n_mapped, n_unmapped, n_acc, n_rej], dtype=np.int64)
                out[p + "accepted_keys"] = np.array(list(acc))
                out[p + "decision_len"] = np.array([len(reads_decision[r]) for r in rb.seqs], dtype=np.int64)
                out[p + "approx_ccl"] = exp.rl_dist.approx_ccl.copy()
                for cname, c in exp.contigs_filt.items():
                    out[f"{p}{cname}_strat"] = np.packbits(c.strat.ravel())
                    out[f"{p}{cname}_coverage_sha"] = np.array(hashlib.sha256(np.ascontiguousarray(c.coverage).tobytes()).hexdigest())
                    out[f"{p}{cname}_scores_sum"] = np.float64(c.scores.sum())
                print(p, "mapped/unmapped/accepted/rejected", n_mapped, n_unmapped, n_acc, n_rej,
                      "accept fraction now", float(np.mean([c.strat.mean() for c in exp.contigs_filt.values()])))
        finally:
            os.chdir(cwd)
    path = GOLDEN / "sim_flow_zymo.npz"
    np.savez_compressed(path, **out)
    print(path.name, f"{path.stat().st_size / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
    main_flow()
