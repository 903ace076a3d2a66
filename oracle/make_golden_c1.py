"""ORACLE / TEST INFRASTRUCTURE — BASELINE.json config 1 at its stated size, recorded from the UPSTREAM simulator itself.

Runs only in the build container (needs /root/reference). Upstream's own `BossRunsSim` (boss/runs/simulation.py) is
constructed from a TOML-equivalent config and driven exactly like `boss.BOSS:main` drives it (BOSS.py:48-56):

    exp.init_sim(); while exp.batch < maxb: exp.process_batch_sim(exp.process_batch_runs_sim)

on the whole of data/BOSS_test_data: `zymo.fa` (9 contigs >= 100 kb, 31 012 581 sites, one contig below 100 kb dropped
at load), `ERR3152366_10k.fq` with `paf_full` / `paf_trunc`, ploidy 1, `bucket_threshold = 0` (as upstream's
tests/config/boss_ch20_sim.toml), sampler seed 1 without shuffling (sampler.py:61,150-158), in two runs:

    A  batchsize = 4000, maxb = 1   the configuration BASELINE.json names (the sampler needs batchsize * (maxb + 1) < 10 000
                                    reads, sampler.py:163-167, so 4000-read batches allow exactly one update)
    B  batchsize = 1000, maxb = 8   eight consecutive updates on the same reads (state evolution: decisions follow the masks)

What the sampler handed out (read ids, both PAF texts) and what the reference computed after every update (decision
counts, read-length staircase, threshold, every contig's mask, a digest of its counters, the sum of its scores) go to
tests/golden/c1_full.npz together with the inputs, so the GPU box needs neither the reference nor this script.

    python -m oracle.make_golden_c1
"""
from __future__ import annotations

import hashlib
import os
import sys
import tempfile
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
REFERENCE = Path(os.environ.get("BOSS_REFERENCE", "/root/reference"))
sys.dont_write_bytecode = True
sys.path[:0] = [str(REPO), str(REPO / "oracle" / "shims"), str(REFERENCE)]

import numpy as np  # noqa: E402

GOLDEN = REPO / "tests" / "golden"
DATA = REFERENCE / "data" / "BOSS_test_data"
RUNS = {"A": dict(batchsize=4000, maxb=1), "B": dict(batchsize=1000, maxb=8)}
_LUT = np.full(256, 255, dtype=np.uint8)
_LUT[np.frombuffer(b"ACGT", dtype=np.uint8)] = np.arange(4, dtype=np.uint8)


def pack2bit(text: str) -> np.ndarray:
    codes = _LUT[np.frombuffer(text.encode(), dtype=np.uint8)]
    assert (codes < 4).all(), "sequence holds characters outside ACGT"
    return np.packbits(np.unpackbits(codes[:, None], axis=1)[:, 6:].ravel())


def run_one(tag: str, batchsize: int, maxb: int, work: Path, out: dict, pool: dict):
    from boss.config import Config
    from boss.runs.simulation import BossRunsSim
    for f in os.listdir(DATA):
        if f.startswith(("zymo.fa", "ERR3152366_10k")) and not (work / f).exists():
            os.symlink(DATA / f, work / f)
    (work / "zymo.mmi").touch()
    cwd = os.getcwd()
    os.chdir(work)
    try:
        args = Config().args
        args.general.name = f"c1{tag}"
        args.general.ref = str(work / "zymo.fa")
        args.general.mmi = str(work / "zymo.mmi")
        args.optional.ploidy = 1
        args.optional.bucket_threshold = 0
        args.simulation.fq = str(work / "ERR3152366_10k.fq")
        args.simulation.paf_full = str(work / "ERR3152366_10k.paf")
        args.simulation.paf_trunc = str(work / "ERR3152366_10k_trunc.paf")
        args.simulation.batchsize, args.simulation.maxb = batchsize, maxb
        exp = BossRunsSim(args)
        exp.init_sim()
        sampled, captured = [], {}
        sample = exp.sampler.sample

        def spy_sample():
            res = sample()
            sampled.append(res)
            return res
        exp.sampler.sample = spy_sample
        find = exp.scoring.find_strat_thread

        def spy_find(benefit, smu, fhat, time_cost):
            strat, thr = find(benefit=benefit, smu=smu, fhat=fhat, time_cost=time_cost)
            captured.update(threshold=float(thr), n_nonzero=int(np.count_nonzero(benefit)), normaliser=float(benefit.max()))
            return strat, thr
        exp.scoring.find_strat_thread = spy_find
        decisions = exp.make_decisions
        counts = {}

        def spy_decisions(**kw):
            res = decisions(**kw)
            counts["c"] = res[2:]
            return res
        exp.make_decisions = spy_decisions
        out[f"{tag}_contigs"] = np.array(list(exp.contigs.keys()))
        out[f"{tag}_n_sites"] = np.int64(exp.ref.n_sites)
        while exp.batch < maxb:
            captured.clear()
            bi = exp.batch
            exp.process_batch_sim(exp.process_batch_runs_sim)
            seqs, quals, bc_names, paf_f, paf_t = sampled[-1]
            for rid, s in seqs.items():
                pool.setdefault(rid, s)
            p = f"{tag}{bi}_"
            out[p + "rids"] = np.array(list(seqs.keys()))
            out[p + "paf_full"] = np.frombuffer(paf_f.encode(), np.uint8)
            out[p + "paf_trunc"] = np.frombuffer(paf_t.encode(), np.uint8)
            out[p + "counts"] = np.array(counts["c"], dtype=np.int64)            # mapped, unmapped, accepted, rejected
            out[p + "approx_ccl"] = exp.rl_dist.approx_ccl.copy()
            out[p + "time_cost"] = np.float64(getattr(exp.rl_dist, "time_cost", np.nan))
            out[p + "updated"] = np.bool_(bool(captured))
            if captured:
                out[p + "threshold"] = np.float64(captured["threshold"])
                out[p + "n_nonzero"] = np.int64(captured["n_nonzero"])
                out[p + "normaliser"] = np.float64(captured["normaliser"])
            for cname, c in exp.contigs_filt.items():
                q = f"{p}{cname}_"
                out[q + "strat"] = np.packbits(c.strat.ravel())
                out[q + "coverage_sha"] = np.array(hashlib.sha256(np.ascontiguousarray(c.coverage).tobytes()).hexdigest())
                out[q + "depth_total"] = np.int64(c.coverage.sum(dtype=np.int64))
                out[q + "scores_sum"] = np.float64(c.scores.sum())
                out[q + "n_dropout"] = np.int64(np.count_nonzero(c.scores == 0.0))
                out[q + "switches"] = np.packbits(c.bucket_switches.ravel())
            print(p, "mapped/unmapped/accepted/rejected", counts["c"], "threshold", captured.get("threshold"),
                  "accept fraction", float(np.mean([c.strat.mean() for c in exp.contigs_filt.values()])), flush=True)
    finally:
        os.chdir(cwd)


def main():
    import logging
    logging.disable(logging.CRITICAL)
    out, pool = {}, {}
    # the reference sequences, 2 bits per base (contigs below 100 kb are kept: the loader has to drop them)
    names, seqs, name, buf = [], {}, None, []
    for line in open(DATA / "zymo.fa"):
        if line.startswith(">"):
            if name is not None:
                seqs[name] = "".join(buf)
            name, buf = line[1:].split()[0], []
            names.append(name)
        else:
            buf.append(line.strip())
    seqs[name] = "".join(buf)
    out["ref_names"] = np.array(names)
    out["ref_lengths"] = np.array([len(seqs[n]) for n in names], dtype=np.int64)
    lut = np.zeros(256, dtype=np.uint8)
    lut[np.frombuffer(b"ACGT", dtype=np.uint8)] = np.arange(4, dtype=np.uint8)
    for n in names:
        raw = np.frombuffer(seqs[n].upper().encode(), dtype=np.uint8)
        codes = lut[raw]                                    # Contig._seq2int: everything outside ACGT is 0 (reference.py:46-68)
        out[f"ref_{n}"] = np.packbits(np.unpackbits(codes[:, None], axis=1)[:, 6:].ravel())
        out[f"ref_{n}_nonacgt"] = np.int64(np.count_nonzero(~np.isin(raw, np.frombuffer(b"ACGT", dtype=np.uint8))))
    for tag, kw in RUNS.items():
        with tempfile.TemporaryDirectory() as td:
            run_one(tag, kw["batchsize"], kw["maxb"], Path(td), out, pool)
        out[f"{tag}_batchsize"], out[f"{tag}_maxb"] = np.int64(kw["batchsize"]), np.int64(kw["maxb"])
    rids = list(pool)
    out["pool_rids"] = np.array(rids)
    out["pool_len"] = np.array([len(pool[r]) for r in rids], dtype=np.int64)
    out["pool_2bit"] = pack2bit("".join(pool[r] for r in rids))
    GOLDEN.mkdir(parents=True, exist_ok=True)
    path = GOLDEN / "c1_full.npz"
    np.savez_compressed(path, **out)
    print(path.name, f"{path.stat().st_size / 1e6:.1f} MB", len(rids), "reads in the pool")


if __name__ == "__main__":
    main()
